/* xfg_stark.h — C ABI of libxfgstark.so, the B200 (sm_100a) proving backend for the XFG burn-mint STARK.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: the path it replaces is the
 * pure-Rust call `air.prove(trace)` at src/burn_mint_prover.rs:124 (winter_prover::Prover::prove on the types bound at
 * src/burn_mint_air.rs:479-531).  A Rust `-sys` crate binds exactly these entry points (rust/xfg-stark-gpu-sys,
 * INTEGRATION.md) and `GpuBurnMintProver: winterfell::Prover` overrides `prove()` with xfg_prove_burn_mint.
 *
 * Conventions: plain pointers and sizes; caller owns every host buffer; field elements cross the boundary as canonical
 * u64 (< p = 2^64 - 2^32 + 1), i.e. `BaseElement::as_int()`; all functions return 0 (XFG_OK) or an XFG_ERR_* code and
 * never throw or abort.  A context is bound to one device and is NOT thread-safe; distinct contexts are independent.
 * There is no CPU fallback: without a usable CUDA device xfg_create fails with XFG_ERR_CUDA.
 */
#ifndef XFG_STARK_H
#define XFG_STARK_H

#include <stddef.h>
#include <stdint.h>
#include "xfg/spec.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xfg_ctx xfg_ctx;

/* replaces winter_air::ProofOptions as built at src/burn_mint_prover.rs:28-35 and accepted by XfgBurnMintProver::with_options
 * (src/burn_mint_prover.rs:44-49); argument meaning and ranges per ProofOptions::new.  Every value ProofOptions::new accepts
 * is served.  The reference's own setting (blowup 8, folding 8, remainder degree >= 7, None / Quadratic) runs on the tuned pipeline; anything
 * else - blowup 2..128, folding 2/4/16, remainder degree 0..3, FieldExtension::Cubic - on the general-options pipeline (same bytes as
 * Winterfell, pinned against proofs of the reference binary; about 1.5x slower at equal work).  Verification covers the same option space.
 * Shapes on which Winterfell itself panics (a FRI layer of one row, an empty remainder) return XFG_ERR_BAD_OPTIONS. */
typedef struct xfg_options {
  uint32_t num_queries;              /* 1..255, smaller than the LDE domain */
  uint32_t blowup_factor;            /* power of two, 2..128 */
  uint32_t grinding_factor;          /* 0..32 */
  uint32_t field_extension;          /* XFG_EXT_NONE = 1, XFG_EXT_QUADRATIC = 2, XFG_EXT_CUBIC = 3 */
  uint32_t fri_folding_factor;       /* 2, 4, 8 or 16 */
  uint32_t fri_remainder_max_degree; /* 2^k - 1, 0..255 */
} xfg_options;

/* replaces BurnMintPublicInputs (src/burn_mint_air.rs:23-71) plus the constants XfgBurnMintAir derives from them and the
 * secret (src/burn_mint_air.rs:124-133, 174-202, 362-365) */
typedef struct xfg_air_consts {
  uint64_t pub_inputs[XFG_NUM_PUB_INPUTS]; /* ToElements order */
  uint64_t txn_hash;                       /* pub.txn_hash as u32 */
  uint64_t recipient_hash;                 /* pub.recipient_hash as u32 */
  uint64_t nullifier;                      /* compute_nullifier(secret) */
  uint64_t commitment;                     /* compute_commitment(secret) */
} xfg_air_consts;

/* device time per winter-prover tracing span, CUDA-event measured */
enum {
  XFG_ST_EXTEND_TRACE = 0, XFG_ST_COMMIT_TRACE, XFG_ST_EVAL_CONSTRAINTS, XFG_ST_COMMIT_CONSTRAINTS, XFG_ST_BUILD_DEEP,
  XFG_ST_EVAL_DEEP, XFG_ST_FRI_LAYERS, XFG_ST_QUERY_POSITIONS, XFG_ST_BUILD_PROOF, XFG_NUM_STAGES
};
typedef struct xfg_stage_times {
  float stage_ms[XFG_NUM_STAGES];
  float h2d_ms;      /* trace upload (0 for the *_device entry point) */
  float device_ms;   /* first kernel .. last kernel, inputs resident */
  float total_ms;    /* upload .. proof material back on the host */
  uint32_t kernel_launches;
  uint64_t h2d_bytes; /* bytes copied host->device inside the call (trace + coin seed) */
  uint64_t d2h_bytes; /* bytes copied device->host inside the call (transcript state + opened rows and paths) */
} xfg_stage_times;

enum {
  XFG_OK = 0,
  XFG_ERR_BAD_ARGS = 1,            /* null pointer, size out of range, non-canonical element */
  XFG_ERR_BAD_OPTIONS = 2,         /* ProofOptions::new range checks (A.2) */
  XFG_ERR_UNSUPPORTED_OPTIONS = 3, /* valid for Winterfell, not implemented: AIR of transition degree > 9, too many assertion steps / live values */
  XFG_ERR_UNSUPPORTED_EXTENSION = 4, /* mirrors ProverError::UnsupportedFieldExtension (kept for the Rust mapping; every extension of f64 is served) */
  XFG_ERR_UNSATISFIED_CONSTRAINT = 5, /* mirrors ProverError::UnsatisfiedTransitionConstraintError / MismatchedConstraintPolynomialDegree */
  XFG_ERR_BUFFER_TOO_SMALL = 6,    /* *out_len holds the required size */
  XFG_ERR_CUDA = 7,                /* CUDA runtime error; see xfg_last_error */
  XFG_ERR_INVALID_INPUT = 8,       /* prove_burn_mint input validation (src/burn_mint_prover.rs:132-208); see xfg_last_error */
  XFG_ERR_TOO_LARGE = 9,           /* n_log2 exceeds the context's max_n_log2 */
  XFG_ERR_INTERNAL = 10
};

/* ---- context ---- */
/* One context = one device + `num_slots` independent proof workspaces (each with its own stream) sized for traces of up to
 * 2^max_n_log2 rows.  Replaces nothing in the reference (its prover is a stateless CPU call). */
int xfg_create(int device, uint32_t max_n_log2, uint32_t num_slots, xfg_ctx** out);
void xfg_destroy(xfg_ctx* ctx);
const char* xfg_strerror(int code);
const char* xfg_last_error(const xfg_ctx* ctx);

/* Per-kernel-family device timing (CUDA events on the launching stream around each launcher call) of proofs run with a
 * non-NULL `times` argument while profiling is on; used by bench.py for the roofline of the dominant kernel.  Replaces
 * nothing in the reference (winter-prover's tracing spans are inert there). */
int xfg_set_profiling(xfg_ctx* ctx, int on);
/* Whole-proof CUDA graphs (default on): proofs run without per-stage timing and without split upload replay a graph captured on
 * first use per (trace length, options, slot), which removes ~50 kernel-launch calls per proof from the host thread. */
int xfg_set_graphs(xfg_ctx* ctx, int on);
int xfg_get_profile(xfg_ctx* ctx, uint32_t cap, uint32_t* count, const char** names, float* ms, uint32_t* launches);

/* ---- whole proof: replaces `air.prove(trace)` (src/burn_mint_prover.rs:124, winter_prover::Prover::prove) ---- */
/* trace: column-major, 7 columns x 2^n_log2 rows (TraceTable::init layout, src/burn_mint_air.rs:475), host memory.
 * out receives StarkProof::to_bytes() (the bytes consumed at src/bin/xfg-stark-cli.rs:533). */
int xfg_prove_burn_mint(xfg_ctx* ctx, const uint64_t* trace_colmajor, uint32_t n_log2, const xfg_air_consts* air,
                        const xfg_options* options, uint8_t* out, size_t out_cap, size_t* out_len, xfg_stage_times* times);
/* same, trace already resident in device memory of ctx's device (d_trace is a device pointer) */
int xfg_prove_burn_mint_device(xfg_ctx* ctx, const uint64_t* d_trace_colmajor, uint32_t n_log2, const xfg_air_consts* air,
                               const xfg_options* options, uint8_t* out, size_t out_cap, size_t* out_len, xfg_stage_times* times);
/* same, the trace given as seven separate column buffers exactly as the reference holds them: `TraceTable::get_column(i)` of the
 * table built at src/burn_mint_air.rs:475 is a `Vec<BaseElement>`, i.e. n u64 words in MONTGOMERY form (x * 2^64 mod p, winter-math 0.8
 * f64::BaseElement).  form = XFG_FORM_MONTGOMERY reads that memory as it is (the factor 2^-64 is folded into the 1/n of the interpolation:
 * no conversion pass, no `as_int()` loop, no copy on the caller's side); XFG_FORM_CANONICAL expects integers < p.  Columns registered with
 * xfg_host_register (or otherwise page-locked) are DMA-ed straight from the caller's memory, pageable columns are staged by the library. */
#define XFG_FORM_CANONICAL  0u
#define XFG_FORM_MONTGOMERY 1u
int xfg_prove_burn_mint_cols(xfg_ctx* ctx, const uint64_t* const cols[XFG_TRACE_WIDTH], uint32_t form, uint32_t n_log2, const xfg_air_consts* air,
                             const xfg_options* options, uint8_t* out, size_t out_cap, size_t* out_len, xfg_stage_times* times);
/* page-locks / releases a caller buffer (cudaHostRegister, read-only + portable) so that traces in it upload without staging */
int xfg_host_register(xfg_ctx* ctx, const void* ptr, size_t bytes);
int xfg_host_unregister(xfg_ctx* ctx, const void* ptr);
/* `count` independent proofs of equal size, pipelined over the context's slots (BASELINE config 4; replaces the sequential
 * loops of examples/winterfell_burn_mint_production.rs:187-195).  traces[i], airs[i] as above; proof i is written at
 * out + i*out_stride (out_stride >= the largest proof) and its length to out_lens[i].
 * Errors: the return value is the FIRST error met; a proof that failed (unsatisfied trace, non-canonical element, ...) or was
 * never started has out_lens[i] == 0, every proof with out_lens[i] > 0 is complete and valid. */
int xfg_prove_burn_mint_batch(xfg_ctx* ctx, uint32_t count, const uint64_t* const* traces, uint32_t n_log2,
                              const xfg_air_consts* airs, const xfg_options* options, uint8_t* out, size_t out_stride,
                              size_t* out_lens, float* total_ms);

/* ---- host-side mirror of XfgBurnMintProver (src/burn_mint_prover.rs) ---- */
/* validate_inputs + secret_to_field_element + compute_recipient_hash + public-input packing (:62-107, :132-221) and the
 * AIR's Keccak scalars (src/burn_mint_air.rs:124-133, 157-202).  Host-only (3 Keccak-256 calls). */
int xfg_burn_mint_pack_inputs(xfg_ctx* ctx, uint64_t burn_amount, uint64_t mint_amount, const uint8_t tx_prefix_hash[32],
                              const uint8_t* recipient_address, size_t recipient_len, const uint8_t* secret, size_t secret_len,
                              uint32_t network_id, uint32_t target_chain_id, uint32_t commitment_version, xfg_air_consts* out);
/* XfgBurnMintAir::build_trace (src/burn_mint_air.rs:442-476) generalised to 2^n_log2 rows (state = floor(4i/n)) */
int xfg_burn_mint_build_trace(const xfg_air_consts* air, uint32_t n_log2, uint64_t* trace_colmajor_out);
/* XfgBurnMintProver::prove_burn_mint (src/burn_mint_prover.rs:62-129), the reference's 8 arguments + the trace length */
int xfg_prove_burn_mint_from_inputs(xfg_ctx* ctx, uint64_t burn_amount, uint64_t mint_amount, const uint8_t tx_prefix_hash[32],
                                    const uint8_t* recipient_address, size_t recipient_len, const uint8_t* secret, size_t secret_len,
                                    uint32_t network_id, uint32_t target_chain_id, uint32_t commitment_version, uint32_t n_log2,
                                    const xfg_options* options, uint8_t* out, size_t out_cap, size_t* out_len, xfg_stage_times* times);

/* ---- generic AIR front-end ----
 * The same pipeline for any main-segment-only AIR whose transition constraints have degree <= 9 and whose assertions are single-point (degree <= 2:
 * ce_blowup = 2 and one composition column, SURVEY.md A.3, on the tuned kernels; degree d = 3..9: d - 1 composition columns and ce_blowup =
 * next_pow2(d - 1), on the general-options pipeline).  The AIR is a straight-line
 * program over the evaluation frame: what the body of `Air::evaluate_transition` computes (e.g. the 4-column XfgBurnAir sketch,
 * src/winterfell_air.rs:87-127, whose constraints are `current[i] - constant`), given as data instead of Rust code.
 * Value ids: [0, width) = frame.current()[i]; [width, 2*width) = frame.next()[i]; [2*width, 2*width + num_constants) =
 * constants[i]; 2*width + num_constants + i = result of code[i] (operands must refer to earlier values).
 * constraint_values[j] is the value id written to result[j].  Assertions are `Assertion::single(column, step, value)`
 * (`Air::get_assertions`, src/winterfell_air.rs:117-124); they are sorted as winter-air does (step, then column) before the
 * boundary coefficients are assigned.  pub_inputs are `PublicInputs::to_elements()`, appended to the coin seed (A.4). */
enum { XFG_OP_ADD = 0, XFG_OP_SUB = 1, XFG_OP_MUL = 2 };
typedef struct xfg_air_instr { uint32_t op, a, b; } xfg_air_instr;
typedef struct xfg_assertion { uint32_t column; uint32_t step; uint64_t value; } xfg_assertion;
typedef struct xfg_air_desc {
  uint32_t width;               /* 1..XFG_AIR_MAX_WIDTH trace columns */
  uint32_t num_pub_inputs;      /* <= XFG_AIR_MAX_PUB_INPUTS */
  uint32_t num_constants;       /* <= XFG_AIR_MAX_CONSTANTS */
  uint32_t num_instr;           /* <= XFG_AIR_MAX_INSTR */
  uint32_t num_constraints;     /* 1..XFG_AIR_MAX_CONSTRAINTS transition constraints */
  uint32_t num_assertions;      /* 1..XFG_AIR_MAX_ASSERTIONS, at most XFG_AIR_MAX_GROUPS distinct steps */
  const uint64_t* pub_inputs;
  const uint64_t* constants;
  const xfg_air_instr* code;
  const uint32_t* constraint_values;
  const xfg_assertion* assertions;
} xfg_air_desc;
/* limits: XFG_AIR_MAX_* in xfg/spec.h */
/* like xfg_create, with workspaces sized for traces of up to max_width columns */
int xfg_create_ex(int device, uint32_t max_n_log2, uint32_t num_slots, uint32_t max_width, xfg_ctx** out);
/* replaces `air.prove(trace)` (winter_prover::Prover::prove) for the AIR described by `air`; trace: column-major, air->width
 * columns x 2^n_log2 rows, host memory (or device memory for the *_device variant).  Errors: XFG_ERR_BAD_ARGS for a malformed
 * description, XFG_ERR_UNSUPPORTED_OPTIONS for degree > 9 / too many groups / too many live values (degrees 3..9 = 2..8 constraint composition
 * columns run on the general-options pipeline),
 * XFG_ERR_UNSATISFIED_CONSTRAINT when the trace violates a constraint or an assertion. */
int xfg_prove_air(xfg_ctx* ctx, const xfg_air_desc* air, const uint64_t* trace_colmajor, uint32_t n_log2, const xfg_options* options,
                  uint8_t* out, size_t out_cap, size_t* out_len, xfg_stage_times* times);
int xfg_prove_air_device(xfg_ctx* ctx, const xfg_air_desc* air, const uint64_t* d_trace_colmajor, uint32_t n_log2, const xfg_options* options,
                         uint8_t* out, size_t out_cap, size_t* out_len, xfg_stage_times* times);

/* `count` independent proofs of equal trace length (descriptions may differ: other constants, assertion values, public inputs, even other
 * programs), pipelined over the context's slots like xfg_prove_burn_mint_batch; traces[i]: host memory, airs[i].width columns x 2^n_log2 rows */
int xfg_prove_air_batch(xfg_ctx* ctx, uint32_t count, const xfg_air_desc* airs, const uint64_t* const* traces, uint32_t n_log2,
                        const xfg_options* options, uint8_t* out, size_t out_stride, size_t* out_lens, float* total_ms);

/* host-only (no device needed): validates and compiles `air` exactly as xfg_prove_air does and reports the compiled size; with cur / next / out
 * non-NULL it also runs the COMPILED register program on that one frame (canonical elements; out = num_constraints results).  A self-test of the
 * front-end's compiler for CPU-only boxes; the proving path never evaluates constraints on the host. */
int xfg_air_compile_check(const xfg_air_desc* air, uint32_t n_log2, uint32_t* num_instr, uint32_t* num_slots, uint32_t* num_groups,
                          const uint64_t* cur, const uint64_t* next, uint64_t* out);

/* ---- one wide trace sharded over the GPUs of a box (BASELINE config 5) ----
 * Replaces DefaultTraceLde::new (src/burn_mint_air.rs:513: interpolate_columns + evaluate_polys_over + commit_to_rows +
 * MerkleTree::new) for a W-column x 2^n_log2-row trace: rank r interpolates and extends columns [r*W/G, (r+1)*W/G); the last
 * NTT pass stores each LDE element directly into the receive buffer of the rank that owns its row (peer stores over NVLink,
 * the all-to-all is fused into the kernel); after a cross-rank barrier every rank hashes its rows [r*N/G, (r+1)*N/G) and
 * returns its subtree root; the caller all-gathers the G roots and finishes the top levels with xfg_merkle_root. */
typedef struct xfg_wide xfg_wide;
int xfg_wide_create(xfg_ctx* ctx, uint32_t n_log2, uint32_t total_cols, uint32_t num_ranks, uint32_t rank, xfg_wide** out);
void xfg_wide_destroy(xfg_wide* w);
void* xfg_wide_recv_ptr(xfg_wide* w);                                   /* device pointer of this rank's receive buffer */
int xfg_wide_ipc_handle(xfg_wide* w, uint8_t out[64]);                  /* cudaIpcMemHandle_t of the receive buffer */
int xfg_wide_open_peers(xfg_wide* w, const uint8_t* handles);           /* num_ranks x 64 bytes (one process per GPU) */
int xfg_wide_set_peer_ptrs(xfg_wide* w, void* const* ptrs);             /* all ranks in one process */
int xfg_wide_extend(xfg_wide* w, const uint64_t* d_cols_local, float* device_ms);   /* device pointer, W/G x n column-major */
int xfg_wide_commit(xfg_wide* w, uint8_t subtree_root[32], float* device_ms);       /* call after a barrier across ranks */
int xfg_wide_read_recv(xfg_wide* w, uint64_t* out);                     /* test hook: W x 8 x n/G elements */

/* ---- batch verification ----
 * Replaces the sequential loop of BatchBurnMintVerifier (src/burn_mint_verifier.rs:371-408) around
 * XfgBurnMintVerifier::verify_with_winterfell -> winterfell::verify::<XfgBurnMintAir, Blake3_256, DefaultRandomCoin>
 * (src/burn_mint_verifier.rs:265-283): one thread block per proof replays the transcript, rebuilds every batch Merkle root,
 * recomputes the DEEP composition at the queried positions and checks the FRI folds and the remainder.
 * results[i] = XFG_VERIFY_OK or the first failing check, named after winterfell::VerifierError / winter_fri::VerifierError.
 * The return value is XFG_OK when the batch ran, whatever the verdicts.  `acceptable` plays the role of AcceptableOptions::OptionSet
 * with one entry (src/burn_mint_verifier.rs:270-276); air[i] are the public inputs and AIR constants of proof i. */
enum {
  XFG_VERIFY_OK = 0,
  XFG_VERIFY_MALFORMED = 1,                        /* ProofDeserializationError: lengths, layout, non-canonical field element */
  XFG_VERIFY_UNACCEPTABLE_OPTIONS = 2,             /* UnacceptableProofOptions */
  XFG_VERIFY_INCONSISTENT_OOD = 3,                 /* InconsistentOodConstraintEvaluations */
  XFG_VERIFY_POW_FAILED = 4,                       /* QuerySeedProofOfWorkVerificationFailed */
  XFG_VERIFY_NUM_QUERIES_MISMATCH = 5,             /* proof's num_unique_queries differs from the drawn positions */
  XFG_VERIFY_TRACE_QUERY_MISMATCH = 6,             /* TraceQueryDoesNotMatchCommitment */
  XFG_VERIFY_CONSTRAINT_QUERY_MISMATCH = 7,        /* ConstraintQueryDoesNotMatchCommitment */
  XFG_VERIFY_FRI_LAYER_COMMITMENT_MISMATCH = 8,    /* FriVerifierError::LayerCommitmentMismatch */
  XFG_VERIFY_FRI_INVALID_LAYER_FOLDING = 9,        /* FriVerifierError::InvalidLayerFolding */
  XFG_VERIFY_FRI_REMAINDER_COMMITMENT_MISMATCH = 10,
  XFG_VERIFY_FRI_REMAINDER_DEGREE_MISMATCH = 11,
  XFG_VERIFY_FRI_INVALID_REMAINDER_FOLDING = 12,
  XFG_VERIFY_FRI_DEGREE_TRUNCATION = 13,
  XFG_VERIFY_COIN = 14                             /* RandomCoinError::FailedToDrawFieldElement */
};
typedef struct xfg_verify_times {
  float host_parse_ms; /* walking the length prefixes + staging copy */
  float h2d_ms;        /* proof bytes to the device */
  float kernel_ms;     /* verify_kernel */
  float total_ms;      /* whole call, wall clock */
  uint64_t h2d_bytes, d2h_bytes;
} xfg_verify_times;
int xfg_verify_burn_mint_batch(xfg_ctx* ctx, uint32_t count, const uint8_t* const* proofs, const size_t* proof_lens,
                               const xfg_air_consts* air /* count entries */, const xfg_options* acceptable,
                               int32_t* results /* count entries */, xfg_verify_times* times /* optional */);
/* the same for proofs of AIRs given as data (the verification counterpart of xfg_prove_air): any ProofOptions, transition degree <= 9; airs: count entries */
int xfg_verify_air_batch(xfg_ctx* ctx, uint32_t count, const uint8_t* const* proofs, const size_t* proof_lens, const xfg_air_desc* airs,
                         const xfg_options* acceptable, int32_t* results /* count entries */, xfg_verify_times* times /* optional */);
const char* xfg_verify_strerror(int code);

/* ---- stage entry points (kernel-level parity tests; host buffers in and out) ---- */
/* `batch` transforms of 2^n_log2 points, contiguous; inverse != 0 = fft::interpolate_poly, else forward evaluation */
int xfg_ntt(xfg_ctx* ctx, uint64_t* data, uint32_t n_log2, uint32_t batch, int inverse);
/* DefaultTraceLde::new (src/burn_mint_air.rs:513): interpolate `cols` (1, 2 or 7) columns, extend by 8 over the coset 7*<w_N>,
 * commit to rows.  lde_out (optional) = cols x 8n evaluations in natural order; root_out = 32 bytes */
int xfg_lde_commit(xfg_ctx* ctx, const uint64_t* cols_colmajor, uint32_t n_log2, uint32_t cols, uint64_t* lde_out, uint8_t root_out[32]);
/* MerkleTree::new over `count` (power of two >= 2) 32-byte leaves; nodes_out (optional) = count node slots as Winterfell */
int xfg_merkle_root(xfg_ctx* ctx, const uint8_t* leaves, size_t count, uint8_t root_out[32], uint8_t* nodes_out);
/* DefaultConstraintEvaluator::evaluate over the 2n-point constraint domain.  lde: 7 x 8n natural order; coeffs: 7 transition
 * then 8 boundary coefficients, `ext` limbs each; out: 2n x ext limbs (element-major), natural order of the CE domain */
int xfg_eval_constraints(xfg_ctx* ctx, const uint64_t* lde, uint32_t n_log2, const xfg_air_consts* air, uint32_t ext,
                         const uint64_t* coeffs, uint64_t* out);
/* folding::apply_drp with folding factor 8 and domain offset 7: evals = 2^nl_log2 x ext limbs (element-major, natural order),
 * alpha = ext limbs; out = 2^(nl_log2-3) x ext limbs */
int xfg_fri_fold_layer(xfg_ctx* ctx, const uint64_t* evals, uint32_t nl_log2, uint32_t ext, const uint64_t* alpha, uint64_t* out);
/* element-wise field self-test of the device arithmetic (winter-math f64::BaseElement): op 0 mul, 1 weak mul,
 * 2 weak add, 3 weak sub, 4 add, 5 sub, 6 inv, 7/8 weak +- b*2^32, 9 un-reduced dot product, 10 weak inversion, 100+S multiply by 2^S; out[i] = canonical result */
int xfg_field_selftest(xfg_ctx* ctx, uint32_t op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out);
/* measured peak of the 32-bit integer ALU pipe (IADD3 / LOP3 / SHF mix, no memory traffic), 1e9 operations per second: the roofline
 * denominator of the BLAKE3 kernels; replaces nothing in the reference */
int xfg_int_pipe_peak(xfg_ctx* ctx, double* gops);
/* the same loop with other instruction mixes (1: IMAD only, 2: 2 ALU + 2 IMAD, 3: 3 ALU + 1 IMAD, 4: IMAD.WIDE only, 5: 3 ALU + 1 IMAD.WIDE, 6: 3-input adds):
 * thread-level instructions per second (1e9).  Measurement tool behind DESIGN.md section 4 (what the FMA pipe can take off the ALU pipe) */
int xfg_pipe_probe(xfg_ctx* ctx, int mode, double* gops);
/* hash_elements of `rows` rows of `limbs` (1, 2, 7, 8 or 16) canonical u64 each, row-major; out = rows x 32 bytes */
int xfg_hash_rows(xfg_ctx* ctx, const uint64_t* rows_rowmajor, size_t rows, uint32_t limbs, uint8_t* out);

/* ---- debug: workspace guard zones.  Every region of a proof's workspace is followed by a guard zone no kernel may write; _fill paints slot 0's
 * slab, _check counts damaged guard words after proofs of the given shape (memory-safety evidence where compute-sanitizer is unavailable) ---- */
int xfg_debug_guard_fill(xfg_ctx* ctx);
int xfg_debug_guard_check(xfg_ctx* ctx, uint32_t n_log2, uint32_t field_extension, uint32_t width, uint32_t fri_remainder_max_degree,
                          uint64_t* violations, int32_t* first_region);
int xfg_debug_poke_guard(xfg_ctx* ctx, uint32_t n_log2);   /* self-test of the checker: damages the first guard zone */

#ifdef __cplusplus
}
#endif
#endif /* XFG_STARK_H */
