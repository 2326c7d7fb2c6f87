//! `GpuBurnMintProver`: a drop-in for the proving half of `XfgBurnMintProver` (src/burn_mint_prover.rs) that keeps the
//! reference's `Prover`/`Air` surface and sends `prove()` to the B200 backend through `xfg-stark-gpu-sys`.
//!
//! NOT COMPILED in this repository's build image (no cargo/rustc); see INTEGRATION.md.
use std::ffi::CStr;
use std::ptr;

use winter_math::StarkField;
use winterfell::{
    crypto::{hashers::Blake3_256, DefaultRandomCoin},
    math::{fields::f64::BaseElement, FieldElement, ToElements},
    matrix::ColMatrix, AuxTraceRandElements, ConstraintCompositionCoefficients, DefaultConstraintEvaluator, DefaultTraceLde, FieldExtension,
    ProofOptions, Prover, ProverError, StarkDomain, StarkProof, Trace, TraceInfo, TracePolyTable, TraceTable,
};
use xfg_stark::burn_mint_air::{BurnMintPublicInputs, XfgBurnMintAir};
use xfg_stark_gpu_sys as sys;

/// Owns one `xfg_ctx` (one GPU).  Not `Sync`: one context per thread / per GPU, as the C ABI requires.
pub struct GpuContext { raw: *mut sys::xfg_ctx }

impl GpuContext {
    pub fn new(device: i32, max_trace_log2: u32, slots: u32) -> Result<Self, String> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::xfg_create(device, max_trace_log2, slots, &mut raw) };
        if rc != sys::XFG_OK { return Err(format!("xfg_create: {}", unsafe { CStr::from_ptr(sys::xfg_strerror(rc)) }.to_string_lossy())); }
        Ok(Self { raw })
    }
    fn last_error(&self) -> String { unsafe { CStr::from_ptr(sys::xfg_last_error(self.raw)) }.to_string_lossy().into_owned() }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { sys::xfg_destroy(self.raw) } } }

fn options_to_c(o: &ProofOptions) -> sys::xfg_options {
    sys::xfg_options {
        num_queries: o.num_queries() as u32,
        blowup_factor: o.blowup_factor() as u32,
        grinding_factor: o.grinding_factor(),
        field_extension: match o.field_extension() { FieldExtension::None => 1, FieldExtension::Quadratic => 2, FieldExtension::Cubic => 3 },
        fri_folding_factor: o.to_fri_options().folding_factor() as u32,
        fri_remainder_max_degree: o.to_fri_options().remainder_max_degree() as u32,
    }
}

/// Same associated types as `impl Prover for XfgBurnMintAir` (src/burn_mint_air.rs:479-531); only `prove` is replaced.
pub struct GpuBurnMintProver {
    pub ctx: GpuContext,
    pub public_inputs: BurnMintPublicInputs,
    /// nullifier / commitment scalars computed by the AIR from the caller's secret (src/burn_mint_air.rs:124-133, 174-202)
    pub nullifier: BaseElement,
    pub commitment: BaseElement,
    pub options: ProofOptions,
}

impl Prover for GpuBurnMintProver {
    type BaseField = BaseElement;
    type Air = XfgBurnMintAir;
    type Trace = TraceTable<BaseElement>;
    type HashFn = Blake3_256<BaseElement>;
    type RandomCoin = DefaultRandomCoin<Self::HashFn>;
    type TraceLde<E: FieldElement<BaseField = BaseElement>> = DefaultTraceLde<E, Self::HashFn>;
    type ConstraintEvaluator<'a, E: FieldElement<BaseField = BaseElement>> = DefaultConstraintEvaluator<'a, XfgBurnMintAir, E>;

    fn get_pub_inputs(&self, _trace: &Self::Trace) -> BurnMintPublicInputs { self.public_inputs.clone() }
    fn options(&self) -> &ProofOptions { &self.options }
    fn new_trace_lde<E: FieldElement<BaseField = BaseElement>>(&self, info: &TraceInfo, main: &ColMatrix<BaseElement>, domain: &StarkDomain<BaseElement>)
        -> (Self::TraceLde<E>, TracePolyTable<E>) { DefaultTraceLde::new(info, main, domain) }
    fn new_evaluator<'a, E: FieldElement<BaseField = BaseElement>>(&self, air: &'a XfgBurnMintAir, aux: AuxTraceRandElements<E>,
        coeffs: ConstraintCompositionCoefficients<E>) -> Self::ConstraintEvaluator<'a, E> { DefaultConstraintEvaluator::new(air, aux, coeffs) }

    /// Whole-proof override: trace columns -> canonical u64 (`as_int`) -> xfg_prove_burn_mint -> `StarkProof::from_bytes`.
    fn prove(&self, trace: Self::Trace) -> Result<StarkProof, ProverError> {
        let n = trace.length();
        let mut cols: Vec<u64> = Vec::with_capacity(7 * n);
        for c in 0..7 { cols.extend(trace.get_column(c).iter().map(|e| e.as_int())); }
        let pi = self.public_inputs.to_elements();
        let mut air = sys::xfg_air_consts::default();
        for (d, s) in air.pub_inputs.iter_mut().zip(pi.iter()) { *d = s.as_int(); }
        air.txn_hash = self.public_inputs.txn_hash.as_int() as u32 as u64;             // src/burn_mint_air.rs:362
        air.recipient_hash = self.public_inputs.recipient_hash.as_int() as u32 as u64; // :365
        air.nullifier = self.nullifier.as_int();
        air.commitment = self.commitment.as_int();
        let opts = options_to_c(&self.options);
        let mut out = vec![0u8; 1 << 20];
        let mut len = 0usize;
        let rc = unsafe { sys::xfg_prove_burn_mint(self.ctx.raw, cols.as_ptr(), n.trailing_zeros(), &air, &opts, out.as_mut_ptr(), out.len(), &mut len, ptr::null_mut()) };
        match rc {
            sys::XFG_OK => { out.truncate(len); StarkProof::from_bytes(&out).map_err(|_| ProverError::UnsupportedFieldExtension(0)) }
            sys::XFG_ERR_UNSATISFIED_CONSTRAINT => Err(ProverError::UnsatisfiedTransitionConstraintError(0)),
            sys::XFG_ERR_UNSUPPORTED_EXTENSION => Err(ProverError::UnsupportedFieldExtension(3)),
            _ => panic!("xfg_prove_burn_mint failed ({rc}): {}", self.ctx.last_error()),
        }
    }
}
