//! Byte-for-byte parity of the GPU backend against real Winterfell 0.8.3, to be run wherever cargo + the crates + a B200 exist:
//!   XFGSTARK_LIB_DIR=.../xfg-stark_b200 cargo test --release -- --nocapture
//! It closes the "parity unpinned" gap of this repository's oracle.  NOT COMPILED here.
//!
//! The reference's own `Air::new` declares 6 constraints / 6 assertions (src/burn_mint_air.rs:309-318) and cannot prove
//!; the CPU side below therefore uses `NormalisedBurnMintAir`, the 7-constraint / 8-assertion AIR of
//! Appendix B.2, which is what the GPU backend and the oracle implement.
use winterfell::{math::fields::f64::BaseElement, FieldExtension, ProofOptions, Prover};

mod normalised_air;   // B.2 AIR + default CPU `Prover` impl + synthetic cases (tests/normalised_air/mod.rs)

#[test]
fn gpu_bytes_equal_winterfell_bytes() {
    for (log_n, ext) in [(6u32, FieldExtension::None), (10, FieldExtension::Quadratic), (16, FieldExtension::None)] {
        let options = ProofOptions::new(42, 8, 4, ext, 8, 31);
        let case = normalised_air::synthetic_case(1usize << log_n, 0, options.clone());   // SplitMix64("XFGSTARK" + index) inputs, §8d
        let cpu = case.cpu_prover.prove(case.trace.clone()).expect("cpu prove").to_bytes();
        let gpu = case.gpu_prover.prove(case.trace.clone()).expect("gpu prove").to_bytes();
        assert_eq!(cpu, gpu, "log_n = {log_n}");
        winterfell::verify::<normalised_air::NormalisedBurnMintAir, winterfell::crypto::hashers::Blake3_256<BaseElement>,
            winterfell::crypto::DefaultRandomCoin<winterfell::crypto::hashers::Blake3_256<BaseElement>>>(
            winterfell::StarkProof::from_bytes(&gpu).unwrap(), case.public_inputs.clone(), &winterfell::AcceptableOptions::OptionSet(vec![options])).expect("verify");
    }
}
