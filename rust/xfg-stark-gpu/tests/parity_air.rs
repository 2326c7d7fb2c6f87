//! Byte-for-byte parity of the generic AIR front-end (`xfg_prove_air`) against real Winterfell 0.8.3 on the 4-register `XfgBurnAir`
//! sketch of the reference (src/winterfell_air.rs:87-127), to be run wherever cargo + the crates + a B200 exist.  NOT COMPILED here.
use winterfell::{
    crypto::{hashers::Blake3_256, DefaultRandomCoin},
    math::{fields::f64::BaseElement, FieldElement},
    matrix::ColMatrix, Air, AirContext, Assertion, AuxTraceRandElements, ConstraintCompositionCoefficients, DefaultConstraintEvaluator, DefaultTraceLde,
    EvaluationFrame, FieldExtension, ProofOptions, Prover, StarkDomain, TraceInfo, TracePolyTable, TraceTable, TransitionConstraintDegree,
};
use xfg_stark_gpu::{xfg_burn_air_program, GpuContext};

/// `XfgBurnAir` (src/winterfell_air.rs:36-127) with its row (commitment, nullifier, amount, network_id) as the statement
#[derive(Clone)]
pub struct Sketch { pub vals: [BaseElement; 4] }
impl winterfell::math::ToElements<BaseElement> for Sketch { fn to_elements(&self) -> Vec<BaseElement> { vec![] } }   // PublicInputs = () in the reference

pub struct SketchAir { context: AirContext<BaseElement>, s: Sketch }
impl Air for SketchAir {
    type BaseField = BaseElement;
    type PublicInputs = Sketch;
    fn new(info: TraceInfo, s: Sketch, options: ProofOptions) -> Self {
        Self { context: AirContext::new(info, vec![TransitionConstraintDegree::new(1); 4], 4, options), s }
    }
    fn context(&self) -> &AirContext<BaseElement> { &self.context }
    fn evaluate_transition<E: FieldElement<BaseField = BaseElement>>(&self, frame: &EvaluationFrame<E>, _p: &[E], r: &mut [E]) {
        for i in 0..4 { r[i] = frame.current()[i] - E::from(self.s.vals[i]); }          // :104-113
    }
    fn get_assertions(&self) -> Vec<Assertion<BaseElement>> { (0..4).map(|i| Assertion::single(i, 0, self.s.vals[i])).collect() }   // :117-124
}
struct CpuProver { options: ProofOptions, s: Sketch }
impl Prover for CpuProver {
    type BaseField = BaseElement;
    type Air = SketchAir;
    type Trace = TraceTable<BaseElement>;
    type HashFn = Blake3_256<BaseElement>;
    type RandomCoin = DefaultRandomCoin<Self::HashFn>;
    type TraceLde<E: FieldElement<BaseField = BaseElement>> = DefaultTraceLde<E, Self::HashFn>;
    type ConstraintEvaluator<'a, E: FieldElement<BaseField = BaseElement>> = DefaultConstraintEvaluator<'a, SketchAir, E>;
    fn get_pub_inputs(&self, _t: &Self::Trace) -> Sketch { self.s.clone() }
    fn options(&self) -> &ProofOptions { &self.options }
    fn new_trace_lde<E: FieldElement<BaseField = BaseElement>>(&self, info: &TraceInfo, main: &ColMatrix<BaseElement>, domain: &StarkDomain<BaseElement>)
        -> (Self::TraceLde<E>, TracePolyTable<E>) { DefaultTraceLde::new(info, main, domain) }
    fn new_evaluator<'a, E: FieldElement<BaseField = BaseElement>>(&self, air: &'a SketchAir, aux: AuxTraceRandElements<E>,
        coeffs: ConstraintCompositionCoefficients<E>) -> Self::ConstraintEvaluator<'a, E> { DefaultConstraintEvaluator::new(air, aux, coeffs) }
}

#[test]
fn generic_front_end_bytes_equal_winterfell_bytes() {
    let vals = [0x1234_5678_90AB_CDEFu64, 987_654_321, 8_000_000, 4].map(BaseElement::new);
    for (log_n, ext) in [(6u32, FieldExtension::None), (10, FieldExtension::Quadratic)] {
        let n = 1usize << log_n;
        let options = ProofOptions::new(42, 8, 4, ext, 8, 31);
        let trace = TraceTable::init((0..4).map(|c| vec![vals[c]; n]).collect());       // generate_execution_trace, src/winterfell_air.rs:186-203 (normalised row)
        let cpu = CpuProver { options: options.clone(), s: Sketch { vals } }.prove(trace).expect("cpu prove").to_bytes();
        let ctx = GpuContext::with_width(0, log_n, 1, 4).expect("B200 context");
        let cols: Vec<u64> = (0..4).flat_map(|c| std::iter::repeat(vals[c].as_int()).take(n)).collect();
        let gpu = xfg_burn_air_program(vals[0], vals[1], vals[2], vals[3]).prove(&ctx, &cols, log_n, &options).expect("gpu prove").to_bytes();
        assert_eq!(cpu, gpu, "log_n = {log_n}");
    }
}
