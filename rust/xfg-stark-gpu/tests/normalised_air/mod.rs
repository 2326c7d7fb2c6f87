//! The normalised BurnMintAir of SURVEY.md Appendix B.2 written against winterfell 0.8, plus the CPU `Prover` with the reference's
//! bindings (src/burn_mint_air.rs:479-531) and the synthetic case generator of SURVEY.md §8(d).  NOT COMPILED in this repository's
//! build image (no cargo/rustc); it is the CPU half of tests/parity.rs.
//!
//! Differences from the reference's `XfgBurnMintAir` are exactly the defect fixes of Appendix B.1: 7 declared constraints / 8 declared
//! assertions, the caller's secret-derived nullifier / commitment (carried beside the 12 public-input elements), trace length n.
use winterfell::{
    crypto::{hashers::Blake3_256, DefaultRandomCoin},
    math::{fields::f64::BaseElement, FieldElement, ToElements},
    matrix::ColMatrix, Air, AirContext, Assertion, AuxTraceRandElements, ConstraintCompositionCoefficients, DefaultConstraintEvaluator,
    DefaultTraceLde, EvaluationFrame, ProofOptions, Prover, StarkDomain, Trace, TraceInfo, TracePolyTable, TraceTable, TransitionConstraintDegree,
};
use xfg_stark_gpu::{GpuBurnMintProver, GpuContext};

#[derive(Clone, Debug)]
pub struct NormalisedPublicInputs {
    pub elements: [BaseElement; 12],   // order of src/burn_mint_air.rs:54-71
    pub nullifier: BaseElement,        // src/burn_mint_air.rs:124-133 with the caller's secret
    pub commitment: BaseElement,       // src/burn_mint_air.rs:174-202
}
impl ToElements<BaseElement> for NormalisedPublicInputs {
    fn to_elements(&self) -> Vec<BaseElement> { self.elements.to_vec() }
}

pub struct NormalisedBurnMintAir { context: AirContext<BaseElement>, pi: NormalisedPublicInputs }

impl Air for NormalisedBurnMintAir {
    type BaseField = BaseElement;
    type PublicInputs = NormalisedPublicInputs;

    fn new(trace_info: TraceInfo, pi: NormalisedPublicInputs, options: ProofOptions) -> Self {
        // r0 and r4 are quadratic (src/burn_mint_air.rs:219, 246), the others linear; 8 assertions (:383-394)
        let d = |k| TransitionConstraintDegree::new(k);
        let degrees = vec![d(2), d(1), d(1), d(1), d(2), d(1), d(1)];
        Self { context: AirContext::new(trace_info, degrees, 8, options), pi }
    }
    fn context(&self) -> &AirContext<BaseElement> { &self.context }

    fn evaluate_transition<E: FieldElement<BaseField = BaseElement>>(&self, frame: &EvaluationFrame<E>, _periodic: &[E], r: &mut [E]) {
        let (c, n) = (frame.current(), frame.next());
        let std_burn = E::from(8_000_000u32);
        let large_burn = std_burn * E::from(1000u32);
        let low32 = |e: BaseElement| E::from(BaseElement::new(e.as_int() as u32 as u64));
        r[0] = (c[0] - std_burn) * (c[0] - large_burn);          // :207-219
        r[1] = c[1] - c[0];                                      // :231
        r[2] = c[2] - low32(self.pi.elements[2]);                // :362
        r[3] = c[3] - low32(self.pi.elements[3]);                // :365
        let d = n[4] - c[4];
        r[4] = d * (d - E::ONE);                                 // :240-246
        r[5] = c[5] - E::from(self.pi.nullifier);                // :264-267
        r[6] = c[6] - E::from(self.pi.commitment);               // :376-377
    }

    fn get_assertions(&self) -> Vec<Assertion<BaseElement>> {
        let last = self.trace_length() - 1;
        let e = &self.pi.elements;
        vec![
            Assertion::single(0, 0, e[0]), Assertion::single(1, 0, e[1]), Assertion::single(2, 0, e[2]), Assertion::single(3, 0, e[3]),
            Assertion::single(4, 0, BaseElement::ZERO), Assertion::single(5, 0, self.pi.nullifier), Assertion::single(6, 0, self.pi.commitment),
            Assertion::single(4, last, BaseElement::new(3)),      // :393 with the literal 63 replaced by n - 1
        ]
    }
}

/// CPU prover with the reference's associated types (src/burn_mint_air.rs:479-531)
pub struct CpuProver { pub options: ProofOptions, pub pi: NormalisedPublicInputs }
impl Prover for CpuProver {
    type BaseField = BaseElement;
    type Air = NormalisedBurnMintAir;
    type Trace = TraceTable<BaseElement>;
    type HashFn = Blake3_256<BaseElement>;
    type RandomCoin = DefaultRandomCoin<Self::HashFn>;
    type TraceLde<E: FieldElement<BaseField = BaseElement>> = DefaultTraceLde<E, Self::HashFn>;
    type ConstraintEvaluator<'a, E: FieldElement<BaseField = BaseElement>> = DefaultConstraintEvaluator<'a, NormalisedBurnMintAir, E>;
    fn get_pub_inputs(&self, _t: &Self::Trace) -> NormalisedPublicInputs { self.pi.clone() }
    fn options(&self) -> &ProofOptions { &self.options }
    fn new_trace_lde<E: FieldElement<BaseField = BaseElement>>(&self, info: &TraceInfo, main: &ColMatrix<BaseElement>, domain: &StarkDomain<BaseElement>)
        -> (Self::TraceLde<E>, TracePolyTable<E>) { DefaultTraceLde::new(info, main, domain) }
    fn new_evaluator<'a, E: FieldElement<BaseField = BaseElement>>(&self, air: &'a NormalisedBurnMintAir, aux: AuxTraceRandElements<E>,
        coeffs: ConstraintCompositionCoefficients<E>) -> Self::ConstraintEvaluator<'a, E> { DefaultConstraintEvaluator::new(air, aux, coeffs) }
}

pub struct Case { pub trace: TraceTable<BaseElement>, pub public_inputs: NormalisedPublicInputs, pub cpu_prover: CpuProver, pub gpu_prover: GpuBurnMintProver }

fn splitmix64(state: &mut u64) -> u64 {
    *state = state.wrapping_add(0x9E3779B97F4A7C15);
    let mut z = *state;
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
    z ^ (z >> 31)
}

/// SplitMix64("XFGSTARK" + index) -> tx_prefix_hash[32], recipient[20], secret[32]; network 4, chain 42161, version 1.
/// The packing of the 12 public inputs and the two Keccak scalars goes through the library's own host mirror
/// (`xfg_burn_mint_pack_inputs`, src/burn_mint_prover.rs:74-107) so that both provers see identical statements.
pub fn synthetic_case(n: usize, index: u64, options: ProofOptions) -> Case {
    let mut st = 0x5846_4753_5441_524Bu64.wrapping_add(index);
    let raw: Vec<u8> = (0..11).flat_map(|_| splitmix64(&mut st).to_le_bytes()).collect();
    let (txp, rcpt, secret) = (&raw[0..32], &raw[32..52], &raw[56..88]);
    let ctx = GpuContext::new(0, n.trailing_zeros().max(3), 1).expect("B200 context");
    let air = ctx.pack_inputs(8_000_000, 8_000_000, txp, rcpt, secret, 4, 42161, 1).expect("pack_inputs");
    let mut elements = [BaseElement::ZERO; 12];
    for (d, s) in elements.iter_mut().zip(air.pub_inputs.iter()) { *d = BaseElement::new(*s); }
    let pi = NormalisedPublicInputs { elements, nullifier: BaseElement::new(air.nullifier), commitment: BaseElement::new(air.commitment) };
    // src/burn_mint_air.rs:442-476 generalised to n rows (state = floor(4 i / n))
    let cols: Vec<Vec<BaseElement>> = (0..7).map(|c| (0..n).map(|i| match c {
        0 => elements[0], 1 => elements[1], 2 => elements[2], 3 => elements[3],
        4 => BaseElement::new((4 * i / n) as u64), 5 => pi.nullifier, _ => pi.commitment }).collect()).collect();
    let trace = TraceTable::init(cols);
    let gpu_prover = GpuBurnMintProver::from_parts(ctx, &air, options.clone());
    Case { trace, public_inputs: pi.clone(), cpu_prover: CpuProver { options, pi }, gpu_prover }
}
