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
    /// workspaces for traces of up to `max_width` columns (generic AIR front-end); `new` sizes them for the 7-column burn-mint trace
    pub fn with_width(device: i32, max_trace_log2: u32, slots: u32, max_width: u32) -> Result<Self, String> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::xfg_create_ex(device, max_trace_log2, slots, max_width, &mut raw) };
        if rc != sys::XFG_OK { return Err(format!("xfg_create_ex: {}", unsafe { CStr::from_ptr(sys::xfg_strerror(rc)) }.to_string_lossy())); }
        Ok(Self { raw })
    }
    pub fn new(device: i32, max_trace_log2: u32, slots: u32) -> Result<Self, String> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { sys::xfg_create(device, max_trace_log2, slots, &mut raw) };
        if rc != sys::XFG_OK { return Err(format!("xfg_create: {}", unsafe { CStr::from_ptr(sys::xfg_strerror(rc)) }.to_string_lossy())); }
        Ok(Self { raw })
    }
    /// `XfgBurnMintProver`'s input half (validation, public-input packing, Keccak scalars: src/burn_mint_prover.rs:74-107, 132-221) through
    /// the library's host mirror `xfg_burn_mint_pack_inputs`
    #[allow(clippy::too_many_arguments)]
    pub fn pack_inputs(&self, burn: u64, mint: u64, tx_prefix_hash: &[u8], recipient: &[u8], secret: &[u8], network_id: u32, target_chain_id: u32,
                       commitment_version: u32) -> Result<sys::xfg_air_consts, String> {
        assert_eq!(tx_prefix_hash.len(), 32);
        let mut air = sys::xfg_air_consts::default();
        let rc = unsafe { sys::xfg_burn_mint_pack_inputs(self.raw, burn, mint, tx_prefix_hash.as_ptr(), recipient.as_ptr(), recipient.len(), secret.as_ptr(),
                                                         secret.len(), network_id, target_chain_id, commitment_version, &mut air) };
        if rc != sys::XFG_OK { return Err(self.last_error()); }
        Ok(air)
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

impl GpuBurnMintProver {
    /// from the packed statement (`GpuContext::pack_inputs`): the 12 public-input elements in `ToElements` order plus the two AIR scalars
    pub fn from_parts(ctx: GpuContext, air: &sys::xfg_air_consts, options: ProofOptions) -> Self {
        let e = |i: usize| BaseElement::new(air.pub_inputs[i]);
        let public_inputs = BurnMintPublicInputs {
            burn_amount: e(0), mint_amount: e(1), txn_hash: e(2), recipient_hash: e(3), state: e(4), tx_prefix_hash_0: e(5), tx_prefix_hash_1: e(6),
            tx_prefix_hash_2: e(7), tx_prefix_hash_3: e(8), network_id: e(9), target_chain_id: e(10), commitment_version: e(11),
        };
        Self { ctx, public_inputs, nullifier: BaseElement::new(air.nullifier), commitment: BaseElement::new(air.commitment), options }
    }
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

    /// Whole-proof override: the seven `TraceTable` columns are handed to the library where they lie - `BaseElement` is a transparent
    /// `u64` holding the Montgomery form x * 2^64 mod p (winter-math 0.8 f64), which `xfg_prove_burn_mint_cols` reads as it is
    /// (XFG_FORM_MONTGOMERY): no `as_int()` pass, no `collect()`, no copy on this side.  -> `StarkProof::from_bytes`.
    fn prove(&self, trace: Self::Trace) -> Result<StarkProof, ProverError> {
        let n = trace.length();
        let cols: [*const u64; 7] = core::array::from_fn(|c| trace.get_column(c).as_ptr() as *const u64);
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
        let rc = unsafe { sys::xfg_prove_burn_mint_cols(self.ctx.raw, cols.as_ptr(), sys::XFG_FORM_MONTGOMERY, n.trailing_zeros(), &air, &opts, out.as_mut_ptr(), out.len(), &mut len, ptr::null_mut()) };
        match rc {
            sys::XFG_OK => { out.truncate(len); StarkProof::from_bytes(&out).map_err(|_| ProverError::UnsupportedFieldExtension(0)) }
            sys::XFG_ERR_UNSATISFIED_CONSTRAINT => Err(ProverError::UnsatisfiedTransitionConstraintError(0)),
            sys::XFG_ERR_UNSUPPORTED_EXTENSION => Err(ProverError::UnsupportedFieldExtension(3)),
            _ => panic!("xfg_prove_burn_mint failed ({rc}): {}", self.ctx.last_error()),
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Generic AIR front-end: any main-segment AIR with transition degree <= 9 and single-point assertions,
// e.g. the 4-register `XfgBurnAir` sketch (src/winterfell_air.rs:87-127).  The AIR's `evaluate_transition` body is
// recorded once as a straight-line program; `prove` then goes through `xfg_prove_air`.
// ------------------------------------------------------------------------------------------------------------------
/// A value of the program: frame.current()[i], frame.next()[i], a constant, or the result of an earlier instruction.
#[derive(Clone, Copy)]
pub enum Val { Cur(u32), Next(u32), Const(u32), Instr(u32) }

#[derive(Default)]
pub struct AirProgram {
    pub width: u32,
    pub pub_inputs: Vec<u64>,
    constants: Vec<u64>,
    code: Vec<(u32, Val, Val)>,
    outputs: Vec<Val>,
    assertions: Vec<sys::xfg_assertion>,
}

impl AirProgram {
    pub fn new(width: u32, pub_inputs: &[BaseElement]) -> Self { Self { width, pub_inputs: pub_inputs.iter().map(|e| e.as_int()).collect(), ..Default::default() } }
    pub fn constant(&mut self, v: BaseElement) -> Val { self.constants.push(v.as_int()); Val::Const(self.constants.len() as u32 - 1) }
    fn emit(&mut self, op: u32, a: Val, b: Val) -> Val { self.code.push((op, a, b)); Val::Instr(self.code.len() as u32 - 1) }
    pub fn add(&mut self, a: Val, b: Val) -> Val { self.emit(sys::XFG_OP_ADD, a, b) }
    pub fn sub(&mut self, a: Val, b: Val) -> Val { self.emit(sys::XFG_OP_SUB, a, b) }
    pub fn mul(&mut self, a: Val, b: Val) -> Val { self.emit(sys::XFG_OP_MUL, a, b) }
    /// `result[j] = v` for the next j
    pub fn constraint(&mut self, v: Val) { self.outputs.push(v) }
    /// `Assertion::single(column, step, value)`
    pub fn assert_single(&mut self, column: usize, step: usize, value: BaseElement) {
        self.assertions.push(sys::xfg_assertion { column: column as u32, step: step as u32, value: value.as_int() })
    }
    fn id(&self, v: Val) -> u32 {
        let (w, c) = (self.width, self.constants.len() as u32);
        match v { Val::Cur(i) => i, Val::Next(i) => w + i, Val::Const(i) => 2 * w + i, Val::Instr(i) => 2 * w + c + i }
    }
    /// trace: column-major canonical u64, `width` columns of 2^n_log2 rows
    pub fn prove(&self, ctx: &GpuContext, trace_colmajor: &[u64], n_log2: u32, options: &ProofOptions) -> Result<StarkProof, ProverError> {
        let code: Vec<sys::xfg_air_instr> = self.code.iter().map(|&(op, a, b)| sys::xfg_air_instr { op, a: self.id(a), b: self.id(b) }).collect();
        let outs: Vec<u32> = self.outputs.iter().map(|&v| self.id(v)).collect();
        let desc = sys::xfg_air_desc {
            width: self.width, num_pub_inputs: self.pub_inputs.len() as u32, num_constants: self.constants.len() as u32, num_instr: code.len() as u32,
            num_constraints: outs.len() as u32, num_assertions: self.assertions.len() as u32, pub_inputs: self.pub_inputs.as_ptr(),
            constants: self.constants.as_ptr(), code: code.as_ptr(), constraint_values: outs.as_ptr(), assertions: self.assertions.as_ptr(),
        };
        let opts = options_to_c(options);
        let mut out = vec![0u8; 1 << 20];
        let mut len = 0usize;
        let rc = unsafe { sys::xfg_prove_air(ctx.raw, &desc, trace_colmajor.as_ptr(), n_log2, &opts, out.as_mut_ptr(), out.len(), &mut len, ptr::null_mut()) };
        match rc {
            sys::XFG_OK => { out.truncate(len); StarkProof::from_bytes(&out).map_err(|_| ProverError::UnsupportedFieldExtension(0)) }
            sys::XFG_ERR_UNSATISFIED_CONSTRAINT => Err(ProverError::UnsatisfiedTransitionConstraintError(0)),
            sys::XFG_ERR_UNSUPPORTED_EXTENSION => Err(ProverError::UnsupportedFieldExtension(3)),
            _ => panic!("xfg_prove_air failed ({rc}): {}", ctx.last_error()),
        }
    }
}

/// The XfgBurnAir sketch (src/winterfell_air.rs:87-127) as a program: constraint i = current[i] - expected_i, assertions at step 0.
pub fn xfg_burn_air_program(commitment: BaseElement, nullifier: BaseElement, amount: BaseElement, network_id: BaseElement) -> AirProgram {
    let mut p = AirProgram::new(4, &[]);
    for (i, v) in [commitment, nullifier, amount, network_id].into_iter().enumerate() {
        let c = p.constant(v);
        let d = p.sub(Val::Cur(i as u32), c);
        p.constraint(d);
        p.assert_single(i, 0, v);
    }
    p
}
