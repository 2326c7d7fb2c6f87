// prover.cu — context, per-size plans, the proof pipeline and proof-byte assembly behind the C ABI (include/xfg_stark.h).
//
// Replaces winter-prover 0.8.3 `Prover::prove` / `generate_proof` as reached from `air.prove(trace)`
// (src/burn_mint_prover.rs:124; types bound at src/burn_mint_air.rs:479-531).  Stage order, transcript order and the wire
// format follow SURVEY.md §3.1 and Appendix A.4-A.12.  Everything between the trace upload and the final copy of the
// opened rows / authentication nodes runs on the device as one dependent chain of launches on the slot's stream; the host
// only serialises (`StarkProof::to_bytes`, `BatchMerkleProof::serialize_nodes`).
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <thread>
#include <tuple>
#include <vector>
#include "../../include/xfg_stark.h"
#include "burn_mint_host.hpp"
#include "proof_bytes.hpp"
#include "air_compile.hpp"
#include "generic_air.cuh"
#include "general.cuh"
#include "general_verify_host.hpp"
#include "merkle.cuh"
#include "ntt.cuh"
#include "stark_kernels.cuh"
#include "transcript.cuh"
#include "verify.cuh"
#include "launch.cuh"

using namespace xfg;

namespace {

// Split upload of a large trace: one copy + event per column, the column's NTTs start as soon as it has landed.  A one-column launch
// fills the GPU for 3.5 waves only (and its interpolation for less than one), so consecutive columns run on UPLOAD_STREAMS different
// streams and fill each other's partial waves (measured e2e at 2^20, ms per proof: one stream with {1,1,2,3}-column groups 4.95, per
// column on 1 / 2 / 3 / 4 streams 5.03 / 4.88 / 4.87 / 4.84).
constexpr int UPLOAD_STREAMS = 4;
constexpr size_t GUARD_WORDS = 64;
constexpr size_t MATERIAL_WORDS = size_t(1) << 20;   // 8 MiB: opened rows + per-position authentication paths
constexpr u32 MIN_LOG = 3, MAX_LOG = 24;

struct DevTable { u64* lo = nullptr; u64* hi = nullptr; u32 nhi = 0; };

struct Plan {
  u32 ln = 0, lN = 0; size_t n = 0, N = 0;
  u32 rem_max_deg = 0, num_layers = 0, layer_log[MAX_LAYERS + 1] = {0};   // layer_log[l] = log2 |domain of FRI layer l|
  u32 rem_log = 0, rem_len = 0;
  u64* slab = nullptr;              // all tables of this plan
  u64* direct = nullptr;            // NttTables::d_* storage (optional)
  NttTables ntt{};                  // tw_* are context-wide
  PowTable wN_inv{};
  u64 *pre_lo = nullptr, *pre_hi = nullptr; u32 pre_hi_stride = 0;   // s_k = 7 w_N^k, k < 8
  u64 *un_lo = nullptr, *un_hi = nullptr; u32 un_hi_stride = 0;      // 7^-1, (7 w_2n)^-1
  u64* d_sk = nullptr;
  u64 s_k[8] = {0}, zinv0 = 0, zinv1 = 0, g_n = 0, g_last = 0, n_inv = 0, inv2 = 0, rem_ninv = 0;
  FriConsts fc{};
};

struct Carve {   // device pointers of one proof, carved from the slot slab for the actual n and extension degree
  u64 *trace_in, *trace_coef, *lde, *ce_evals, *ce_tmp, *h_coef, *h_lde, *deep, *rem_in, *rem_coef;
  Digest *trace_tree, *comp_tree;
  u64* fri_evals[MAX_LAYERS + 1]; Digest* fri_tree[MAX_LAYERS];
  size_t words;
};

struct GraphKey {      // everything that is baked into the captured launch sequence (kernel arguments, grids, copy sizes)
  const void* plan; const void* trace; int D; u32 q, g; u32 width, seed_count, prog_instr;   // prog_instr: 0 = burn-mint kernels, else generic program length + 1
  std::array<const void*, XFG_TRACE_WIDTH> host_src;                                         // split upload: the column copies are part of the graph
  bool fill;                                                                                 // trace built on the device from the AIR constants
  u32 ncoef;                                                                                 // constraint + assertion count (baked into the trace-root step of the tree kernel)
  bool operator<(const GraphKey& o) const {
    return std::tie(plan, trace, D, q, g, width, seed_count, prog_instr, host_src, fill, ncoef) < std::tie(o.plan, o.trace, o.D, o.q, o.g, o.width, o.seed_count, o.prog_instr, o.host_src, o.fill, o.ncoef);
  }
};
struct GraphEntry { cudaGraphExec_t exec; unsigned launches; };
struct GoGraphKey {    // what a general-options launch sequence bakes in (general_api.inc)
  const void* plan; const void* base; const void* trace; int D; u32 W, K, num_assertions, ncoef, q, g; u64 in_scale;
  bool operator<(const GoGraphKey& o) const { return std::tie(plan, base, trace, D, W, K, num_assertions, ncoef, q, g, in_scale) < std::tie(o.plan, o.base, o.trace, o.D, o.W, o.K, o.num_assertions, o.ncoef, o.q, o.g, o.in_scale); }
};
struct ProfRec { const char* name; size_t e0, e1; unsigned launches; };
struct GoPlanHolder { GoPlan plan; std::vector<void*> allocs; };   // a general-options plan and the device tables it owns (general_api.inc)

struct Slot {
  cudaStream_t st = nullptr, copy_st = nullptr, aux_st[3] = {nullptr, nullptr, nullptr};   // copy_st: column-wise trace upload overlapped with the first NTTs; aux_st: every other column group
  cudaEvent_t col_ev[XFG_TRACE_WIDTH] = {nullptr}, fork_ev = nullptr, fork2_ev = nullptr, join_ev[3] = {nullptr, nullptr, nullptr};
  u32 tail_threads = 1024;                                  // block size of the fused FRI tail kernel: 1024 for one proof at a time, 256 when proofs are pipelined over the slots
  bool fill_trace = false;                                  // burn-mint trace generated on the device (xfg_prove_burn_mint_from_inputs): no upload
  bool split_upload = false; const u64* up_cols[XFG_TRACE_WIDTH] = {nullptr};   // split upload: the (pinned) host columns the copies read, issued inside enqueue_proof
  u64 in_scale = 1;                                         // 1 for canonical input, R^-1 = 2^-64 for Montgomery-form columns: folded into the 1/n of the interpolation
  u64* slab = nullptr; size_t slab_words = 0;
  ProofState* d_state = nullptr; u64* d_partial = nullptr; u64* d_material = nullptr;
  ProofState* h_state = nullptr; u64* h_material = nullptr; u64* h_seed = nullptr; u64* h_trace = nullptr;   // h_seed: pinned mirror of the init block of ProofState
  AirParams* d_air = nullptr; AirParams* h_air = nullptr;
  // generic AIR front-end (xfg_prove_air): compiled program + AIR-sized state; W = trace width of the proof in flight
  GenProgram* d_prog = nullptr; GenProgram* h_prog = nullptr; GenState* d_gen = nullptr; u64 (*h_ood)[2] = nullptr;
  bool generic = false; u32 W = XFG_TRACE_WIDTH, seed_count = 8 + XFG_NUM_PUB_INPUTS;
  GoState* d_go = nullptr; GoState* h_go = nullptr;   // state of a general-options proof (allocated on first use)
  std::map<GoGraphKey, GraphEntry> go_graphs;            // whole-proof CUDA graphs of the general-options pipeline
  u64* go_slab = nullptr; size_t go_slab_words = 0;    // its workspace when the slab above is too small (blowup > 8, cubic extension at the maximum length)
  std::map<GraphKey, GraphEntry> graphs;            // whole-proof CUDA graphs, one per (plan, extension, options, trace pointer)
  cudaEvent_t ev[XFG_NUM_STAGES + 3] = {nullptr};
  // in-flight proof (batch mode)
  bool busy = false; const Plan* plan = nullptr; int D = 1; xfg_options opt{}; u32 proof_index = 0; bool timed = false;
  GatherTasks tasks{}; size_t mat_words = 0;
  // optional per-kernel-family timing (xfg_set_profiling): events around each launcher call on this slot's stream
  std::vector<cudaEvent_t> pev; std::vector<ProfRec> prof; size_t pev_used = 0;
};

}  // namespace

thread_local unsigned g_xfg_launches = 0;

struct VerifyBufs { u8* h = nullptr; u8* d = nullptr; size_t cap = 0; };   // grow-only pinned staging + device copy of a verification batch

struct xfg_ctx {
  int device = 0; u32 max_log = 0, max_width = XFG_TRACE_WIDTH;
  VerifyBufs vbufs;
  std::vector<Slot> slots;
  u64 *tw_fwd = nullptr, *tw_inv = nullptr;
  std::map<u64, Plan> plans;
  std::map<std::array<u32, 4>, GoPlanHolder> go_plans;   // general-options plans (general_api.inc), keyed by (log2 n, blowup, folding, remainder degree)
  std::string last_error;
  bool profiling = false, graphs = true;
  std::vector<std::string> prof_names; std::vector<float> prof_ms; std::vector<unsigned> prof_launches;   // last profiled proof
};

namespace {

#define CU(call)                                                                                         \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                                                  \
      ctx->last_error = std::string(#call) + ": " + cudaGetErrorString(e_); return XFG_ERR_CUDA; } } while (0)

int fail(xfg_ctx* ctx, int code, const std::string& msg) { if (ctx) ctx->last_error = msg; return code; }

std::vector<u64> pow_series(u64 base, size_t count) { std::vector<u64> v(count); u64 x = 1; for (size_t i = 0; i < count; i++) { v[i] = x; x = gl_mul(x, base); } return v; }

// ProofOptions::new range checks (A.2).  Options outside the tuned set (blowup 8, folding 8, remainder degree >= 7, None / Quadratic) are
// served by the general-options pipeline (general_api.inc: go_needed); `tuned_only` callers (the batch verifier) refuse them instead.
bool go_needed(const xfg_options& o);
int prove_general(xfg_ctx* ctx, const xfg_air_desc& air, const xfg_air_consts* bm, const u64* h_trace, const u64* const* h_cols, u32 form, const u64* d_trace, bool fill,
                  u32 n_log2, const xfg_options& o, u8* out, size_t cap, size_t* out_len, xfg_stage_times* times);
int prove_general_burn_mint(xfg_ctx* ctx, const xfg_air_consts& air, const u64* h_trace, const u64* const* h_cols, u32 form, const u64* d_trace, bool fill,
                            u32 n_log2, const xfg_options& o, u8* out, size_t cap, size_t* out_len, xfg_stage_times* times);
int verify_general_burn_mint(xfg_ctx* ctx, u32 count, const u8* const* proofs, const size_t* lens, const xfg_air_consts* air, const xfg_options& o, int32_t* results, xfg_verify_times* times);
int check_options(xfg_ctx* ctx, const xfg_options* o, u32 n_log2, bool tuned_only = false) {
  auto pow2 = [](u32 x) { return x && !(x & (x - 1)); };
  if (o->field_extension != XFG_EXT_NONE && o->field_extension != XFG_EXT_QUADRATIC && o->field_extension != XFG_EXT_CUBIC) return fail(ctx, XFG_ERR_BAD_OPTIONS, "invalid field extension");
  if (o->num_queries < 1) return fail(ctx, XFG_ERR_BAD_OPTIONS, "number of queries must be greater than 0");
  if (o->num_queries > 255) return fail(ctx, XFG_ERR_BAD_OPTIONS, "number of queries cannot be greater than 255");
  if (!pow2(o->blowup_factor)) return fail(ctx, XFG_ERR_BAD_OPTIONS, "blowup factor must be a power of 2");
  if (o->blowup_factor < 2 || o->blowup_factor > 128) return fail(ctx, XFG_ERR_BAD_OPTIONS, "blowup factor out of range");
  if (o->grinding_factor > 32) return fail(ctx, XFG_ERR_BAD_OPTIONS, "grinding factor cannot be greater than 32");
  if (!pow2(o->fri_folding_factor) || o->fri_folding_factor < 2 || o->fri_folding_factor > 16) return fail(ctx, XFG_ERR_BAD_OPTIONS, "FRI folding factor must be a power of 2 in 2..16");
  if (o->fri_remainder_max_degree > 255 || !pow2(o->fri_remainder_max_degree + 1)) return fail(ctx, XFG_ERR_BAD_OPTIONS, "FRI polynomial remainder degree must be one less than a power of two");
  if (tuned_only && o->field_extension == XFG_EXT_CUBIC) return fail(ctx, XFG_ERR_UNSUPPORTED_EXTENSION, "the cubic extension is not supported on this path");
  if (tuned_only && go_needed(*o)) return fail(ctx, XFG_ERR_UNSUPPORTED_OPTIONS, "this path implements blowup factor 8, FRI folding factor 8 and fri_remainder_max_degree >= 7 (the reference's settings)");
  if ((u64)o->num_queries >= ((u64)o->blowup_factor << n_log2)) return fail(ctx, XFG_ERR_BAD_OPTIONS, "number of queries must be smaller than the LDE domain size");
  return XFG_OK;
}

int upload_table(xfg_ctx* ctx, u64* d_lo, u64* d_hi, u32 nhi, u64 base) {
  std::vector<u64> lo = pow_series(base, POW_LO), hi = pow_series(gl_pow(base, POW_LO), nhi);
  CU(cudaMemcpy(d_lo, lo.data(), POW_LO * 8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_hi, hi.data(), (size_t)nhi * 8, cudaMemcpyHostToDevice));
  return XFG_OK;
}

int get_plan(xfg_ctx* ctx, u32 ln, u32 rem_max_deg, const Plan** out) {
  const u64 key = ((u64)ln << 32) | rem_max_deg;
  auto it = ctx->plans.find(key);
  if (it != ctx->plans.end()) { *out = &it->second; return XFG_OK; }
  Plan p; p.ln = ln; p.lN = ln + 3; p.n = size_t(1) << ln; p.N = p.n * 8; p.rem_max_deg = rem_max_deg;
  // FriOptions::num_fri_layers (A.10)
  { size_t dom = p.N, mx = (size_t)(rem_max_deg + 1) * 8; u32 lg = p.lN; p.layer_log[0] = lg;
    while (dom > mx) { dom /= 8; lg -= 3; p.num_layers++; if (p.num_layers > MAX_LAYERS) return fail(ctx, XFG_ERR_INTERNAL, "too many FRI layers"); p.layer_log[p.num_layers] = lg; }
    p.rem_log = lg; p.rem_len = (u32)(dom / 8);
    if (dom < 8 || p.rem_log > NTT_SINGLE_MAX_LOG || p.rem_len > MAX_REMAINDER) return fail(ctx, XFG_ERR_UNSUPPORTED_OPTIONS, "FRI remainder domain out of range"); }
  const u32 nhi_n = (u32)std::max<size_t>(1, p.n >> POW_LO_BITS), nhi_N = (u32)std::max<size_t>(1, p.N >> POW_LO_BITS);
  // slab: wn_fwd, wn_inv (lo+hi each), wN_inv, pre[8], un[2], s_k[8]
  const size_t words = 2 * (POW_LO + nhi_n) + (POW_LO + nhi_N) + 8 * (POW_LO + nhi_n) + 2 * (POW_LO + nhi_n) + 8;
  CU(cudaMalloc(&p.slab, words * 8));
  u64* w = p.slab;
  auto take = [&](size_t k) { u64* r = w; w += k; return r; };
  const u64 wN = gl_root_of_unity(p.lN), wn = gl_root_of_unity(ln);
  u64 *a, *b; int rc;
  a = take(POW_LO); b = take(nhi_n); if ((rc = upload_table(ctx, a, b, nhi_n, wn))) return rc; p.ntt.wn_fwd = PowTable{a, b};
  a = take(POW_LO); b = take(nhi_n); if ((rc = upload_table(ctx, a, b, nhi_n, gl_inv(wn)))) return rc; p.ntt.wn_inv = PowTable{a, b};
  a = take(POW_LO); b = take(nhi_N); if ((rc = upload_table(ctx, a, b, nhi_N, gl_inv(wN)))) return rc; p.wN_inv = PowTable{a, b};
  p.pre_lo = take(8 * POW_LO); p.pre_hi = take(8 * (size_t)nhi_n); p.pre_hi_stride = nhi_n;
  for (u32 k = 0; k < 8; k++) {
    p.s_k[k] = gl_mul(XFG_GENERATOR, gl_pow(wN, k));
    if ((rc = upload_table(ctx, p.pre_lo + (size_t)k * POW_LO, p.pre_hi + (size_t)k * nhi_n, nhi_n, p.s_k[k]))) return rc;
  }
  p.un_lo = take(2 * POW_LO); p.un_hi = take(2 * (size_t)nhi_n); p.un_hi_stride = nhi_n;
  if ((rc = upload_table(ctx, p.un_lo, p.un_hi, nhi_n, gl_inv(p.s_k[0])))) return rc;
  if ((rc = upload_table(ctx, p.un_lo + POW_LO, p.un_hi + nhi_n, nhi_n, gl_inv(p.s_k[4])))) return rc;
  p.d_sk = take(8);
  CU(cudaMemcpy(p.d_sk, p.s_k, 64, cudaMemcpyHostToDevice));
  p.ntt.tw_fwd = ctx->tw_fwd; p.ntt.tw_inv = ctx->tw_inv;
  p.g_n = wn; p.g_last = gl_pow(wn, p.n - 1);
  p.zinv0 = gl_inv(gl_sub(gl_pow(p.s_k[0], p.n), 1)); p.zinv1 = gl_inv(gl_sub(gl_pow(p.s_k[4], p.n), 1));
  p.n_inv = gl_inv((u64)p.n); p.inv2 = gl_inv(2); p.rem_ninv = gl_inv((u64)1 << p.rem_log);
  // full-size twiddle tables for the four-step transforms (88 MB at n = 2^20, 1.4 GB at 2^24; XFG_NTT_DIRECT=0 keeps the two-level lookups, for A/B runs)
  { const char* e = getenv("XFG_NTT_DIRECT"); const size_t dw = (e && e[0] == '0') ? 0 : ntt_direct_words(ln, 8, 2);
    if (dw) {
      if (cudaMalloc(&p.direct, dw * 8) != cudaSuccess) { cudaGetLastError(); p.direct = nullptr; }   // optional: fall back to the lookups
      else {
        ntt_build_direct(p.ntt, ln, p.direct, p.n_inv, p.pre_lo, p.pre_hi, p.pre_hi_stride, 8, p.un_lo, p.un_hi, p.un_hi_stride, 2);
        CU(cudaDeviceSynchronize());
      }
    } }
  // the table uploads above are synchronous copies from pageable memory, which may return before the DMA has landed; the consumers run on
  // non-blocking streams that are not ordered against the legacy stream, so order them here once per plan
  CU(cudaDeviceSynchronize());
  const u64 w8i = gl_inv(gl_root_of_unity(3));
  p.fc.w8i[0] = 1; for (int i = 1; i < 4; i++) p.fc.w8i[i] = gl_mul(p.fc.w8i[i - 1], w8i);
  p.fc.inv8 = gl_inv(8); p.fc.inv7 = gl_inv(XFG_GENERATOR);
  auto ins = ctx->plans.emplace(key, p);
  *out = &ins.first->second; return XFG_OK;
}

size_t slab_words_for(u32 ln, int D, size_t W) {
  const size_t n = size_t(1) << ln, N = 8 * n;
  size_t w = W * n + W * n + W * N + 8 * N + 2 * D * n + 2 * D * n + D * n + D * N + 8 * N + D * N;
  w += (size_t)D * N / 7 + 64 * MAX_LAYERS;   // FRI layer evaluations l >= 1
  w += 8 * N / 7 + 64 * MAX_LAYERS;           // FRI trees
  w += 2 * (size_t)D * 2048 + 64;             // remainder in / coefficients
  w += (size_t)GUARD_WORDS * (16 + 2 * MAX_LAYERS);   // guard zones between the regions (see carve)
  return w;
}

// Every region of the workspace is followed by a guard zone of GUARD_WORDS words that no kernel may touch: compute-sanitizer is not
// available on this GPU pool, so out-of-bounds writes are hunted with xfg_debug_guard_fill / xfg_debug_guard_check (tests/test_gpu_guards.py):
// fill the slab with a pattern, prove, and require every guard zone (and the slack behind the last region) to be intact.
void carve(const Slot& s, const Plan& p, int D, Carve& c, std::vector<std::pair<size_t, size_t>>* guards = nullptr) {
  u64* w = s.slab; const size_t n = p.n, N = p.N, W = s.W;
  auto take = [&](size_t k) { u64* r = w; w += (k + 7) & ~size_t(7); if (guards) guards->push_back({(size_t)(w - s.slab), (size_t)GUARD_WORDS}); w += GUARD_WORDS; return r; };
  c.trace_in = take(W * n); c.trace_coef = take(W * n); c.lde = take(W * N);
  c.trace_tree = reinterpret_cast<Digest*>(take(8 * N));
  c.ce_evals = take(2 * D * n); c.ce_tmp = take(2 * D * n); c.h_coef = take(D * n); c.h_lde = take(D * N);
  c.comp_tree = reinterpret_cast<Digest*>(take(8 * N));
  c.deep = take(D * N);
  c.fri_evals[0] = c.deep;
  for (u32 l = 1; l <= p.num_layers; l++) c.fri_evals[l] = take((size_t)D << p.layer_log[l]);
  for (u32 l = 0; l < p.num_layers; l++) c.fri_tree[l] = reinterpret_cast<Digest*>(take(size_t(1) << p.layer_log[l]));   // 2 * Nl/8 digests
  c.rem_in = take((size_t)D << p.rem_log); c.rem_coef = take((size_t)D << p.rem_log);
  c.words = (size_t)(w - s.slab);
}

// layout of the material buffer for this proof; fills the gather tasks
size_t build_gather(const Plan& p, int D, u32 W, const xfg_options& o, const Carve& c, GatherTasks& g) {
  const u32 q = o.num_queries; size_t off = 0; u32 t = 0;
  auto add = [&](const u64* src, const Digest* tree, u64 limb_stride, u64 coset_n, u64 R, u64 M, u32 J, u32 limbs, u32 depth, int layer) {
    GatherTask& k = g.t[t++]; k.src = src; k.tree = tree; k.limb_stride = limb_stride; k.coset_n = coset_n; k.R = R; k.M = M; k.J = J; k.limbs = limbs;
    k.depth = depth; k.fri_layer = layer; k.rows_off = off; off += ((size_t)q * J * limbs + 3) & ~size_t(3); k.paths_off = off; off += (size_t)q * depth * 4;
  };
  add(c.lde, c.trace_tree, p.N, p.n, 0, p.N, 1, W, p.lN, -1);
  add(c.h_lde, c.comp_tree, p.N, p.n, 0, p.N, 1, (u32)D, p.lN, -1);
  for (u32 l = 0; l < p.num_layers; l++) {
    const u64 Nl = u64(1) << p.layer_log[l], R = Nl / 8;
    add(c.fri_evals[l], c.fri_tree[l], l == 0 ? p.N : Nl, l == 0 ? p.n : 0, R, R, 8, (u32)D, p.layer_log[l] - 3, (int)l);
  }
  g.count = t; return off;
}

// ---- enqueue the whole proof on the slot's stream (no host synchronisation) ----
void prepare_inputs(Slot& s, const Plan& p, const xfg_options& o, const xfg_air_consts& air) {
  seed_elements(p.ln, o, XFG_TRACE_WIDTH, air.pub_inputs, XFG_NUM_PUB_INPUTS, s.h_seed);
  s.generic = false; s.W = XFG_TRACE_WIDTH; s.seed_count = 8 + XFG_NUM_PUB_INPUTS;
  AirParams& ap = *s.h_air;
  ap.txn = air.txn_hash; ap.rcpt = air.recipient_hash; ap.nullifier = air.nullifier; ap.commitment = air.commitment;
  ap.assert0[0] = air.pub_inputs[XFG_PI_BURN]; ap.assert0[1] = air.pub_inputs[XFG_PI_MINT]; ap.assert0[2] = air.pub_inputs[XFG_PI_TXN_HASH];
  ap.assert0[3] = air.pub_inputs[XFG_PI_RECIPIENT_HASH]; ap.assert0[4] = 0; ap.assert0[5] = air.nullifier; ap.assert0[6] = air.commitment;
  ap.g_last = p.g_last;
}

int enqueue_proof(xfg_ctx* ctx, Slot& s, const Plan& p, int D, const xfg_options& o, const u64* d_trace, bool timed) {
  Carve c; carve(s, p, D, c);
  if (c.words > s.slab_words) return fail(ctx, XFG_ERR_TOO_LARGE, "workspace too small for this trace length");
  cudaStream_t st = s.st; const u32 ln = p.ln; const size_t n = p.n, N = p.N;
  int ev = 0;
  auto mark = [&]() { if (timed) cudaEventRecord(s.ev[ev], st); ev++; };
  const u64* trace_src = d_trace ? d_trace : c.trace_in;
  const bool profiling = ctx->profiling && timed;
  s.prof.clear(); s.pev_used = 0;
  // PROF(name, launches...) brackets a launcher call with events when profiling is on
  auto prof_begin = [&](const char* name) { if (!profiling) return; if (s.pev.size() < s.pev_used + 2) { s.pev.resize(s.pev_used + 2, nullptr); }
    for (size_t i = s.pev_used; i < s.pev_used + 2; i++) if (!s.pev[i]) cudaEventCreate(&s.pev[i]);
    cudaEventRecord(s.pev[s.pev_used], st); s.prof.push_back(ProfRec{name, s.pev_used, s.pev_used + 1, g_xfg_launches}); s.pev_used += 2; };
  auto prof_end = [&]() { if (!profiling) return; ProfRec& r = s.prof.back(); cudaEventRecord(s.pev[r.e1], st); r.launches = g_xfg_launches - r.launches; };
#define PROF(name, ...) do { prof_begin(name); __VA_ARGS__; prof_end(); } while (0)

  // per-proof inputs (coin seed elements, AIR constants) were written to pinned memory by prepare_inputs(); copying them here keeps
  // the whole launch sequence replayable as a CUDA graph
  const u32 W = s.W; const bool gen = s.generic;
  if (s.split_upload && !d_trace) {   // one copy + event per column on the copy stream (forked here so that the whole sequence is capturable as a
    CU(cudaEventRecord(s.fork_ev, st)); CU(cudaStreamWaitEvent(s.copy_st, s.fork_ev, 0));      // CUDA graph); a column's NTTs start as soon as it has landed
    for (int g = 0; g < XFG_TRACE_WIDTH; g++) {
      CU(cudaMemcpyAsync(c.trace_in + (size_t)g * n, s.up_cols[g], n * 8, cudaMemcpyHostToDevice, s.copy_st));
      CU(cudaEventRecord(s.col_ev[g], s.copy_st));
    }
  }
  // init block of the proof state (coin seed elements, their count, cleared error flags, unset nonce) in one copy: no seeding kernel.  The pinned
  // mirror is completed by launch_prepared for EVERY proof (a replayed graph reads it at execution time)
  CU(cudaMemcpyAsync(s.d_state, s.h_seed, PROOF_INIT_BYTES, cudaMemcpyHostToDevice, st));
  if (gen) CU(cudaMemcpyAsync(s.d_prog, s.h_prog, offsetof(GenProgram, code) + (size_t)s.h_prog->num_instr * sizeof(GenInstr), cudaMemcpyHostToDevice, st));
  else CU(cudaMemcpyAsync(s.d_air, s.h_air, sizeof(AirParams), cudaMemcpyHostToDevice, st));
  mark();   // ev0: start of device work
  if (s.fill_trace && !d_trace && !gen) PROF("trace_fill", launch_trace_fill(st, c.trace_in, s.d_air, ln));
  // 1 ---- extend_execution_trace: interpolate the 7 columns, evaluate on the 8 cosets s_k * <w_n>
  // With a split upload the trace goes column by column so that a column's NTTs start as soon as its copy has landed.  (Running
  // the HBM-resident path column by column as well - to keep one column's 64 MB four-step intermediate inside the L2 - was
  // measured: 7x smaller grids cost more (5.88 vs 5.46 ms per proof) than the saved DRAM traffic gains on these ALU-bound kernels.)
  // Consecutive column groups alternate between two streams: a group's launches are 1-3 columns wide (3.5 waves of CTAs per column), and
  // the partial last wave of one group is filled by the next group's kernels instead of idling.
  const bool waits = s.split_upload && !d_trace, two = waits && !profiling;
  const int nstreams = two ? UPLOAD_STREAMS : 1;
  if (two) { CU(cudaEventRecord(s.fork2_ev, st)); for (int a = 0; a + 1 < nstreams; a++) CU(cudaStreamWaitEvent(s.aux_st[a], s.fork2_ev, 0)); }
  for (int g = 0; g < (waits ? XFG_TRACE_WIDTH : 1); g++) {
    const int c0 = waits ? g : 0, per = waits ? 1 : (int)W;
    const size_t off = (size_t)c0 * n;
    cudaStream_t gs = (g % nstreams) ? s.aux_st[g % nstreams - 1] : st;
    if (waits) CU(cudaStreamWaitEvent(gs, s.col_ev[g], 0));
    // every trace element must be a canonical field element: checked by the pass of the interpolation that reads the trace
    { NttJob j{}; j.src = trace_src + off; j.dst = c.trace_coef + off; j.ln = ln; j.batch = per; j.src_tstride = n; j.dst_tstride = n; j.src_div = 1;
      j.canon_flag = &s.d_state->error_flags; j.canon_bit = ERR_FLAG_NONCANONICAL;
      j.inverse = true; j.scale = gen ? p.n_inv : gl_mul(p.n_inv, s.in_scale); PROF("ntt.interpolate_trace", ntt_batch(gs, p.ntt, j)); }
    { NttJob j{}; j.src = c.trace_coef + off; j.dst = c.lde + off * 8; j.ln = ln; j.batch = per * 8; j.src_tstride = n; j.dst_tstride = n; j.src_div = 8;
      j.inverse = false; j.scale = 1; j.pre_lo = p.pre_lo; j.pre_hi = p.pre_hi; j.pre_hi_stride = p.pre_hi_stride; PROF("ntt.lde_trace", ntt_batch(gs, p.ntt, j)); }
  }
  for (int a = 0; a + 1 < nstreams; a++) { CU(cudaEventRecord(s.join_ev[a], s.aux_st[a])); CU(cudaStreamWaitEvent(st, s.join_ev[a], 0)); }
  mark();
  //   ---- compute_execution_trace_commitment
  if (W == 1 || W == 2 || W == XFG_TRACE_WIDTH) PROF("commit_rows.trace", launch_commit_rows(st, c.lde, N, (int)W, ln, c.trace_tree));
  else PROF("commit_rows.trace", launch_commit_rows_wide(st, c.lde, N, W, ln, c.trace_tree));
  // commit_trace + the constraint composition coefficients run on the CTA that computes the root (RootStep): no separate transcript launch
  auto ticket = [&](int i) -> unsigned* { return i < 16 ? reinterpret_cast<unsigned*>(reinterpret_cast<char*>(s.d_state) + offsetof(ProofState, tickets)) + i : nullptr; };
  { RootStep rs{}; rs.kind = 1; rs.D = D; rs.ps = s.d_state; rs.ticket = ticket(0);
    if (gen) { rs.out = reinterpret_cast<u64(*)[2]>(reinterpret_cast<char*>(s.d_gen) + offsetof(GenState, coef)); rs.count = s.h_prog->num_constraints + s.h_prog->num_assertions; }
    else { rs.out = reinterpret_cast<u64(*)[2]>(reinterpret_cast<char*>(s.d_state) + offsetof(ProofState, tcoef)); rs.count = XFG_NUM_TRANSITION + XFG_NUM_ASSERTIONS; }
    PROF("tree_upper.trace", merkle_build_upper(st, c.trace_tree, n, &rs)); }
  mark();
  // 2 ---- evaluate_constraints
  // The constraint-evaluation domain is cosets 0 and 4 of the LDE domain, so the evaluations are written where the composition LDE needs them:
  // limb l, coset 4k' of h_lde = transform 2l + k' at stride 4n (no copy after the interpolation has verified the degree)
  if (gen) PROF("constraints", launch_gen_constraints(st, D, c.lde, ln, s.d_prog, s.d_gen, p.ntt.wn_fwd, p.s_k[0], p.s_k[4], p.zinv0, p.zinv1, p.g_last, c.h_lde, 4 * n));
  else PROF("constraints", launch_constraints(st, D, c.lde, ln, s.d_air, s.d_state, p.ntt.wn_fwd, p.s_k[0], p.s_k[4], p.zinv0, p.zinv1, c.h_lde, 4 * n));
  mark();
  // 3 ---- commit_to_constraint_evaluations: coset interpolation (2 cosets of size n), composition column, LDE, commitment
  { NttJob j{}; j.src = c.h_lde; j.dst = c.ce_tmp; j.ln = ln; j.batch = 2 * D; j.src_tstride = 4 * n; j.dst_tstride = n; j.src_div = 1;
    j.inverse = true; j.scale = p.n_inv; j.post_lo = p.un_lo; j.post_hi = p.un_hi; j.post_hi_stride = p.un_hi_stride; j.post_div = 2; PROF("ntt.interpolate_comp", ntt_batch(st, p.ntt, j)); }
  PROF("combine", launch_combine(st, c.ce_tmp, ln, D, p.inv2, c.h_coef, s.d_state));
  // composition LDE: cosets 0 and 4 of the LDE domain ARE the constraint-evaluation domain, and once combine_kernel has verified
  // that the 2n-point interpolant has degree < n the polynomial's values there are the evaluations the constraint kernel already
  // wrote into those slots of h_lde - compute only cosets 1,2,3,5,6,7
  { NttJob j{}; j.src = c.h_coef; j.dst = c.h_lde; j.ln = ln; j.batch = D * 6; j.src_tstride = n; j.dst_tstride = n; j.src_div = 6;
    j.coset_map = 0x765321; j.dst_cosets = 8;
    j.inverse = false; j.scale = 1; j.pre_lo = p.pre_lo; j.pre_hi = p.pre_hi; j.pre_hi_stride = p.pre_hi_stride; PROF("ntt.lde_comp", ntt_batch(st, p.ntt, j)); }
  PROF("commit_rows.comp", launch_commit_rows(st, c.h_lde, N, D, ln, c.comp_tree));
  { RootStep rs{}; rs.kind = 2; rs.D = D; rs.ps = s.d_state; rs.g_n = p.g_n; rs.ticket = ticket(1);      // commit_constraints, draw z, z g
    PROF("tree_upper.comp", merkle_build_upper(st, c.comp_tree, n, &rs)); }
  mark();
  // 4 ---- build_deep_composition_poly: OOD frame + coefficients
  PROF("ood", launch_ood(st, D, c.trace_coef, c.h_coef, ln, W, s.d_state, s.d_partial));
  if (gen) PROF("transcript", launch_gen_ood_finish(st, D, s.d_state, s.d_gen, W, s.d_partial, ood_num_blocks(ln)));
  else PROF("transcript", launch_ood_finish(st, D, s.d_state, s.d_partial, ood_num_blocks(ln)));
  mark();
  // 5 ---- evaluate_deep_composition_poly (pointwise) + leaves of the first FRI layer
  { const u64* dcoef = gen ? &s.d_gen->dcoef[0][0] : reinterpret_cast<const u64*>(reinterpret_cast<const char*>(s.d_state) + offsetof(ProofState, dcoef));
    PROF("deep", launch_deep(st, D, c.lde, c.h_lde, ln, s.d_state, dcoef, W, p.ntt.wn_fwd, p.d_sk, c.deep, p.num_layers ? c.fri_tree[0] : nullptr)); }
  mark();
  // 6 ---- compute_fri_layers: one tree / commit / fold launch set per large layer; every layer of <= 2^13 evaluations (FRI_TAIL_MAX_LOG), the remainder, the
  //   ---- grinding nonce and the query positions (7) run in ONE single-CTA launch (fri_tail.cu)
  u32 first_tail = p.num_layers;
  while (first_tail > 0 && p.layer_log[first_tail - 1] <= FRI_TAIL_MAX_LOG) first_tail--;
  for (u32 l = 0; l < first_tail; l++) {
    { RootStep rs{}; rs.kind = 3; rs.D = D; rs.ps = s.d_state; rs.layer = l; rs.ticket = ticket(2 + (int)l);     // commit_fri_layer, draw alpha
      PROF("fri.tree", merkle_build_upper(st, c.fri_tree[l], size_t(1) << (p.layer_log[l] - 3), &rs)); }
    PROF("fri.fold", launch_fri_fold(st, D, c.fri_evals[l], l == 0 ? N : (size_t(1) << p.layer_log[l]), l == 0, p.layer_log[l], l, s.d_state, p.wN_inv, p.lN, p.fc,
                    c.fri_evals[l + 1], size_t(1) << p.layer_log[l + 1], l + 1 < p.num_layers ? c.fri_tree[l + 1] : nullptr));
  }
  { const u64* rin = c.fri_evals[p.num_layers]; const size_t Rm = size_t(1) << p.rem_log;
    if (p.num_layers == 0) { launch_coset_to_natural(st, c.deep, c.rem_in, ln, D, N, Rm); rin = c.rem_in; }
    FriTailArgs ta{}; ta.first_layer = first_tail; ta.num_layers = p.num_layers;
    for (u32 l = 0; l <= p.num_layers; l++) { ta.layer_log[l] = p.layer_log[l]; ta.evals[l] = c.fri_evals[l]; ta.limb_stride[l] = l == 0 ? N : (u64(1) << p.layer_log[l]); }
    for (u32 l = 0; l < p.num_layers; l++) ta.tree[l] = c.fri_tree[l];
    ta.rem_in = rin; ta.rem_stride = Rm; ta.rem_log = p.rem_log; ta.rem_len = p.rem_len; ta.rem_ninv = p.rem_ninv;
    ta.tw_inv = p.ntt.tw_inv; ta.un_lo = p.un_lo; ta.wN_inv = p.wN_inv; ta.lN = p.lN; ta.fc = p.fc;
    ta.grinding = o.grinding_factor; ta.num_queries = o.num_queries; ta.do_grind = o.grinding_factor <= FRI_TAIL_MAX_GRIND;
    PROF("fri.tail", launch_fri_tail(st, D, ta, s.d_state, s.tail_threads));
    mark();
    // 7 ---- determine_query_positions (inside the tail kernel unless the grinding factor needs the whole chip)
    if (!ta.do_grind) {
      PROF("grind", launch_grind(st, s.d_state, o.grinding_factor));
      PROF("transcript", launch_positions(st, s.d_state, o.num_queries, p.lN, p.num_layers));
    } }
  mark();
  // 8 ---- build_proof_object: gather opened rows + authentication nodes, copy out
  const size_t mat_words = build_gather(p, D, s.W, o, c, s.tasks);
  if (mat_words > MATERIAL_WORDS) return fail(ctx, XFG_ERR_INTERNAL, "material buffer too small");
  s.mat_words = mat_words;
  PROF("gather", launch_gather(st, s.tasks, s.d_state, s.d_material));
  mark();   // end of device work
  CU(cudaMemcpyAsync(s.h_state, s.d_state, sizeof(ProofState), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(s.h_material, s.d_material, mat_words * 8, cudaMemcpyDeviceToHost, st));
  if (gen) CU(cudaMemcpyAsync(s.h_ood, s.d_gen->ood_frame, (size_t)2 * W * 16, cudaMemcpyDeviceToHost, st));
  mark();
  CU(cudaGetLastError());
#undef PROF
  s.busy = true; s.plan = &p; s.D = D; s.opt = o; s.timed = timed;
  return XFG_OK;
}

// Runs the proof's launch sequence: as a cached CUDA graph (captured on first use per plan / extension / options / trace pointer)
// when no per-stage timing is requested and the upload is not split, otherwise launch by launch.
// the slot's per-proof inputs (seed elements, AIR constants or compiled program, width) have been prepared in its pinned mirrors
int launch_prepared(xfg_ctx* ctx, Slot& s, const Plan& p, int D, const xfg_options& o, const u64* d_trace, bool timed) {
  { u32* tail = reinterpret_cast<u32*>(s.h_seed + MAX_SEED_LIMBS); tail[0] = s.seed_count; tail[1] = 0; tail[2] = tail[3] = 0xFFFFFFFFu;    // seed_count, error_flags = 0, nonce = ~0
    for (int i = 0; i < 16; i++) tail[4 + i] = 0; }                                                                                       // tree tickets
  { Carve c; carve(s, p, D, c); if (c.words > s.slab_words) return fail(ctx, XFG_ERR_TOO_LARGE, "workspace too small for this trace length"); }
  const bool use_graph = ctx->graphs && !timed && !ctx->profiling;
  if (!use_graph) return enqueue_proof(ctx, s, p, D, o, d_trace, timed);
  const bool split = s.split_upload && !d_trace;
  std::array<const void*, XFG_TRACE_WIDTH> srcs{}; if (split) for (int c = 0; c < XFG_TRACE_WIDTH; c++) srcs[c] = s.up_cols[c];
  const GraphKey key{&p, d_trace, D, o.num_queries + (s.tail_threads << 16), o.grinding_factor + (s.in_scale != 1 ? 256u : 0u), s.W, s.seed_count, s.generic ? s.h_prog->num_instr + 1 : 0, srcs, s.fill_trace && !d_trace,
                     s.generic ? s.h_prog->num_constraints + s.h_prog->num_assertions : 0u};
  auto it = s.graphs.find(key);
  if (it == s.graphs.end()) {
    if (s.graphs.size() >= 16) { for (auto& kv : s.graphs) cudaGraphExecDestroy(kv.second.exec); s.graphs.clear(); }   // bounded cache (callers that keep changing the device trace pointer)
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(s.st, cudaStreamCaptureModeThreadLocal));
    const unsigned before = g_xfg_launches;
    const int rc = enqueue_proof(ctx, s, p, D, o, d_trace, false);
    const cudaError_t ce = cudaStreamEndCapture(s.st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) { ctx->last_error = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); return XFG_ERR_CUDA; }
    cudaGraphExec_t exec = nullptr;
    CU(cudaGraphInstantiate(&exec, graph, 0));
    cudaGraphDestroy(graph);
    it = s.graphs.emplace(key, GraphEntry{exec, g_xfg_launches - before}).first;
  } else {
    Carve c; carve(s, p, D, c); s.mat_words = build_gather(p, D, s.W, o, c, s.tasks);
    g_xfg_launches += it->second.launches;
    s.busy = true; s.plan = &p; s.D = D; s.opt = o; s.timed = false; s.prof.clear();
  }
  CU(cudaGraphLaunch(it->second.exec, s.st));
  return XFG_OK;
}
int launch_proof(xfg_ctx* ctx, Slot& s, const Plan& p, int D, const xfg_options& o, const xfg_air_consts& air, const u64* d_trace, bool timed) {
  prepare_inputs(s, p, o, air);
  return launch_prepared(ctx, s, p, D, o, d_trace, timed);
}

// ---- host serialisation (byte writer, batch Merkle paths, seed elements: proof_bytes.hpp) ----
// StarkProof::to_bytes (A.12)
void assemble(const Plan& p, int D, u32 W, const u64 (*ood_frame)[2], const xfg_options& o, const ProofState& s, const u64* mat, const GatherTasks& g, std::vector<u8>& bytes) {
  Out out; out.b.reserve(size_t(1) << 18);
  // Context
  out.u8_(W); out.u8_(0); out.u8_(0); out.u8_(p.ln); out.u16_(0); out.u8_(8); out.u64_(XFG_P);
  out.u8_(o.num_queries); out.u8_(o.blowup_factor); out.u8_(o.grinding_factor); out.u8_(o.field_extension); out.u8_(o.fri_folding_factor); out.u8_(o.fri_remainder_max_degree);
  out.u8_(s.num_positions);
  // Commitments
  out.u16_(32 * (3 + p.num_layers));
  out.raw(&s.trace_root, 32); out.raw(&s.constraint_root, 32);
  for (u32 l = 0; l < p.num_layers; l++) out.raw(&s.fri_roots[l], 32);
  out.raw(&s.remainder_commitment, 32);
  // trace + constraint Queries: u32 len + values, u32 len + paths
  auto queries = [&](const GatherTask& t, const u32* pos, u32 cnt) {
    const size_t vbytes = (size_t)cnt * t.J * t.limbs * 8;
    out.u32_(vbytes); out.raw(mat + t.rows_off, vbytes);
    Out pth; batch_paths(pos, cnt, mat + t.paths_off, t.depth, t.M, pth);
    out.u32_(pth.b.size()); out.raw(pth.b.data(), pth.b.size());
  };
  queries(g.t[0], s.positions, s.num_positions);
  queries(g.t[1], s.positions, s.num_positions);
  // OodFrame
  out.u16_(1 + 2 * (size_t)W * D * 8); out.u8_(2);
  for (u32 i = 0; i < 2 * W; i++) for (int l = 0; l < D; l++) out.u64_(ood_frame[i][l]);
  out.u16_(D * 8); for (int l = 0; l < D; l++) out.u64_(s.hz[l]);
  // FriProof
  out.u8_(p.num_layers);
  for (u32 l = 0; l < p.num_layers; l++) queries(g.t[2 + l], s.fri_positions[l], s.fri_num_positions[l]);
  out.u16_((size_t)s.remainder_len * D * 8);
  for (u32 i = 0; i < s.remainder_len; i++) for (int l = 0; l < D; l++) out.u64_(s.remainder[i][l]);
  out.u8_(0);   // FriProof::num_partitions as log2 (one partition), pinned against the reference binary
  out.u64_(s.nonce);
  bytes.swap(out.b);
}

int finish_proof(xfg_ctx* ctx, Slot& s, u8* out, size_t cap, size_t* out_len, xfg_stage_times* times) {
  *out_len = 0;   // every failure below leaves a defined (empty) result for this proof
  CU(cudaStreamSynchronize(s.st));
  s.busy = false;
  const ProofState& hs = *s.h_state;
  if (!s.prof.empty()) {   // aggregate the per-launcher event pairs by name, in first-seen order
    ctx->prof_names.clear(); ctx->prof_ms.clear(); ctx->prof_launches.clear();
    for (const ProfRec& r : s.prof) {
      float ms = 0; cudaEventElapsedTime(&ms, s.pev[r.e0], s.pev[r.e1]);
      size_t k = 0; while (k < ctx->prof_names.size() && ctx->prof_names[k] != r.name) k++;
      if (k == ctx->prof_names.size()) { ctx->prof_names.push_back(r.name); ctx->prof_ms.push_back(0); ctx->prof_launches.push_back(0); }
      ctx->prof_ms[k] += ms; ctx->prof_launches[k] += r.launches;
    }
  }
  if (hs.error_flags & ERR_FLAG_NONCANONICAL) return fail(ctx, XFG_ERR_BAD_ARGS, "non-canonical trace element");
  if (hs.error_flags & ERR_FLAG_DEGREE) return fail(ctx, XFG_ERR_UNSATISFIED_CONSTRAINT, "UnsatisfiedTransitionConstraintError: the trace does not satisfy the AIR (composition polynomial degree too high)");
  if (hs.error_flags & ERR_FLAG_COIN) return fail(ctx, XFG_ERR_INTERNAL, "FailedToDrawFieldElement");
  std::vector<u8> bytes; assemble(*s.plan, s.D, s.W, s.generic ? s.h_ood : hs.ood_frame, s.opt, hs, s.h_material, s.tasks, bytes);
  if (times && s.timed) {
    for (int i = 0; i < XFG_NUM_STAGES; i++) cudaEventElapsedTime(&times->stage_ms[i], s.ev[i], s.ev[i + 1]);
    cudaEventElapsedTime(&times->device_ms, s.ev[0], s.ev[XFG_NUM_STAGES]);
  }
  *out_len = bytes.size();
  if (bytes.size() > cap) return fail(ctx, XFG_ERR_BUFFER_TOO_SMALL, "output buffer too small");
  std::memcpy(out, bytes.data(), bytes.size());
  return XFG_OK;
}

// waits for and discards every in-flight proof (error paths of the batch entry point; also run before any new call so that a
// failed batch can never leave a slot pointing at a caller buffer that no longer exists)
void drain_slots(xfg_ctx* ctx) {
  for (Slot& s : ctx->slots) if (s.busy) { cudaStreamSynchronize(s.st); s.busy = false; }
}

int check_air(xfg_ctx* ctx, const xfg_air_consts* air) {
  for (int i = 0; i < XFG_NUM_PUB_INPUTS; i++) if (air->pub_inputs[i] >= XFG_P) return fail(ctx, XFG_ERR_BAD_ARGS, "non-canonical public input");
  if (air->txn_hash >= XFG_P || air->recipient_hash >= XFG_P || air->nullifier >= XFG_P || air->commitment >= XFG_P) return fail(ctx, XFG_ERR_BAD_ARGS, "non-canonical AIR constant");
  return XFG_OK;
}

// trace upload: straight from the caller's buffers when they are page-locked (cudaHostAlloc / cudaHostRegister / xfg_host_register), otherwise
// staged through the slot's pinned buffer (large columns by one host thread each).  Canonicity (< p) is checked on the device.
// cols[c]: column c of the trace (W columns of n elements; for a contiguous column-major trace cols[c] = trace + c * n).
int upload_cols(xfg_ctx* ctx, Slot& s, const Plan& p, int D, const u64* const* cols, bool allow_split) {
  Carve c; carve(s, p, D, c);
  if (c.words > s.slab_words) return fail(ctx, XFG_ERR_TOO_LARGE, "workspace too small for this trace length");
  const size_t n = p.n, W = s.W;
  const u64* src[XFG_AIR_MAX_WIDTH]; bool contiguous = true, stage_any = false;
  for (size_t k = 0; k < W; k++) {
    cudaPointerAttributes at{}; const bool pinned = cudaPointerGetAttributes(&at, cols[k]) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();   // unregistered host memory may leave a sticky-free error code behind on older drivers
    src[k] = pinned ? cols[k] : nullptr; stage_any |= !pinned;
    if (k && cols[k] != cols[0] + k * n) contiguous = false;
    if (k && pinned != (src[0] != nullptr)) contiguous = false;
  }
  if (stage_any) {
    auto stage = [&](size_t k) { if (!src[k]) { std::memcpy(s.h_trace + k * n, cols[k], n * 8); src[k] = s.h_trace + k * n; } };
    if (n * 8 >= (size_t(1) << 20)) { std::vector<std::thread> th; for (size_t k = 0; k < W; k++) if (!src[k]) th.emplace_back(stage, k); for (auto& t : th) t.join(); }
    else for (size_t k = 0; k < W; k++) stage(k);
  }
  s.split_upload = allow_split && p.ln >= 17 && !s.generic;
  if (s.split_upload) {      // the column copies are issued by enqueue_proof (copy stream forked from the proof's stream), so that the
    for (size_t k = 0; k < W; k++) s.up_cols[k] = src[k];   // whole end-to-end sequence can be captured and replayed as one CUDA graph
  } else if (contiguous) {
    CU(cudaMemcpyAsync(c.trace_in, src[0], W * n * 8, cudaMemcpyHostToDevice, s.st));
  } else {
    for (size_t k = 0; k < W; k++) CU(cudaMemcpyAsync(c.trace_in + k * n, src[k], n * 8, cudaMemcpyHostToDevice, s.st));
  }
  return XFG_OK;
}
int upload_trace(xfg_ctx* ctx, Slot& s, const Plan& p, int D, const u64* h_trace, bool allow_split) {
  const u64* cols[XFG_AIR_MAX_WIDTH];
  for (size_t k = 0; k < s.W; k++) cols[k] = h_trace + k * p.n;
  return upload_cols(ctx, s, p, D, cols, allow_split);
}

int prove_common(xfg_ctx* ctx, const u64* h_trace, const u64* d_trace, u32 n_log2, const xfg_air_consts* air, const xfg_options* o,
                 u8* out, size_t cap, size_t* out_len, xfg_stage_times* times, bool fill = false, const u64* const* h_cols = nullptr, u32 form = XFG_FORM_CANONICAL) {
  if (!ctx || !air || !o || !out_len || (!h_trace && !d_trace && !fill && !h_cols) || (!out && cap)) return fail(ctx, XFG_ERR_BAD_ARGS, "null argument");
  if (form != XFG_FORM_CANONICAL && form != XFG_FORM_MONTGOMERY) return fail(ctx, XFG_ERR_BAD_ARGS, "unknown element form");
  if (h_cols) for (int c = 0; c < XFG_TRACE_WIDTH; c++) if (!h_cols[c]) return fail(ctx, XFG_ERR_BAD_ARGS, "null column");
  if (n_log2 < MIN_LOG || n_log2 > MAX_LOG) return fail(ctx, XFG_ERR_BAD_ARGS, "trace length must be 2^3 .. 2^24");
  if (n_log2 > ctx->max_log) return fail(ctx, XFG_ERR_TOO_LARGE, "trace longer than the context was created for");
  int rc;
  if ((rc = check_options(ctx, o, n_log2)) || (rc = check_air(ctx, air))) return rc;
  if (go_needed(*o)) return prove_general_burn_mint(ctx, *air, h_trace, h_cols, form, d_trace, fill, n_log2, *o, out, cap, out_len, times);   // options outside the tuned 8/8 set
  CU(cudaSetDevice(ctx->device));
  drain_slots(ctx);
  const Plan* p; if ((rc = get_plan(ctx, n_log2, o->fri_remainder_max_degree, &p))) return rc;
  Slot& s = ctx->slots[0];
  const int D = o->field_extension == XFG_EXT_QUADRATIC ? 2 : 1;
  g_xfg_launches = 0;
  if (times) { std::memset(times, 0, sizeof *times); cudaEventRecord(s.ev[XFG_NUM_STAGES + 2], s.st); }
  s.split_upload = false; s.generic = false; s.W = XFG_TRACE_WIDTH; s.fill_trace = fill; s.tail_threads = 1024;
  // Montgomery-form input (x * 2^64 mod p, what winter-math's BaseElement holds in memory): the interpolation is linear, so the factor
  // 2^-64 rides on its 1/n scale (2^64 = 2^32 - 1 mod p) and costs nothing
  s.in_scale = form == XFG_FORM_MONTGOMERY ? gl_inv(0xFFFFFFFFull) : 1;
  if (h_trace && (rc = upload_trace(ctx, s, *p, D, h_trace, true))) return rc;
  if (h_cols && (rc = upload_cols(ctx, s, *p, D, h_cols, true))) return rc;
  if ((rc = launch_proof(ctx, s, *p, D, *o, *air, d_trace, times != nullptr))) return rc;
  rc = finish_proof(ctx, s, out, cap, out_len, times);
  if (times) {
    cudaEventElapsedTime(&times->h2d_ms, s.ev[XFG_NUM_STAGES + 2], s.ev[0]);
    cudaEventElapsedTime(&times->total_ms, s.ev[XFG_NUM_STAGES + 2], s.ev[XFG_NUM_STAGES + 1]);
    times->kernel_launches = g_xfg_launches;
    times->h2d_bytes = ((h_trace || h_cols) ? (size_t)s.W * p->n * 8 : 0) + PROOF_INIT_BYTES;
    times->d2h_bytes = sizeof(ProofState) + s.mat_words * 8;
  }
  return rc;
}

}  // namespace

// =============================================================== C ABI ===============================================================
extern "C" {

const char* xfg_strerror(int code) {
  switch (code) {
    case XFG_OK: return "ok";
    case XFG_ERR_BAD_ARGS: return "bad arguments";
    case XFG_ERR_BAD_OPTIONS: return "invalid proof options";
    case XFG_ERR_UNSUPPORTED_OPTIONS: return "proof options not supported by this backend";
    case XFG_ERR_UNSUPPORTED_EXTENSION: return "UnsupportedFieldExtension";
    case XFG_ERR_UNSATISFIED_CONSTRAINT: return "UnsatisfiedTransitionConstraintError";
    case XFG_ERR_BUFFER_TOO_SMALL: return "output buffer too small";
    case XFG_ERR_CUDA: return "CUDA error";
    case XFG_ERR_INVALID_INPUT: return "invalid burn-mint input";
    case XFG_ERR_TOO_LARGE: return "trace too large for this context";
    default: return "internal error";
  }
}
const char* xfg_verify_strerror(int code) {
  switch (code) {
    case XFG_VERIFY_OK: return "";
    case XFG_VERIFY_MALFORMED: return "ProofDeserializationError";
    case XFG_VERIFY_UNACCEPTABLE_OPTIONS: return "UnacceptableProofOptions";
    case XFG_VERIFY_INCONSISTENT_OOD: return "InconsistentOodConstraintEvaluations";
    case XFG_VERIFY_POW_FAILED: return "QuerySeedProofOfWorkVerificationFailed";
    case XFG_VERIFY_NUM_QUERIES_MISMATCH: return "NumberOfQueriesMismatch";
    case XFG_VERIFY_TRACE_QUERY_MISMATCH: return "TraceQueryDoesNotMatchCommitment";
    case XFG_VERIFY_CONSTRAINT_QUERY_MISMATCH: return "ConstraintQueryDoesNotMatchCommitment";
    case XFG_VERIFY_FRI_LAYER_COMMITMENT_MISMATCH: return "LayerCommitmentMismatch";
    case XFG_VERIFY_FRI_INVALID_LAYER_FOLDING: return "InvalidLayerFolding";
    case XFG_VERIFY_FRI_REMAINDER_COMMITMENT_MISMATCH: return "RemainderCommitmentMismatch";
    case XFG_VERIFY_FRI_REMAINDER_DEGREE_MISMATCH: return "RemainderDegreeMismatch";
    case XFG_VERIFY_FRI_INVALID_REMAINDER_FOLDING: return "InvalidRemainderFolding";
    case XFG_VERIFY_FRI_DEGREE_TRUNCATION: return "DegreeTruncation";
    case XFG_VERIFY_COIN: return "FailedToDrawFieldElement";
    default: return "unknown verification error";
  }
}
const char* xfg_last_error(const xfg_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int xfg_create(int device, uint32_t max_n_log2, uint32_t num_slots, xfg_ctx** out) { return xfg_create_ex(device, max_n_log2, num_slots, XFG_TRACE_WIDTH, out); }
int xfg_create_ex(int device, uint32_t max_n_log2, uint32_t num_slots, uint32_t max_width, xfg_ctx** out) {
  if (!out || max_n_log2 < MIN_LOG || max_n_log2 > MAX_LOG || num_slots < 1 || num_slots > 64 || max_width < 1 || max_width > XFG_AIR_MAX_WIDTH) return XFG_ERR_BAD_ARGS;
  if (max_width < XFG_TRACE_WIDTH) max_width = XFG_TRACE_WIDTH;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return XFG_ERR_CUDA;   // no CPU fallback
  xfg_ctx* ctx = new xfg_ctx; ctx->device = device; ctx->max_log = max_n_log2; ctx->max_width = max_width;
  auto bail = [&](int rc) { xfg_destroy(ctx); return rc; };
#define CUB(call) do { if ((call) != cudaSuccess) return bail(XFG_ERR_CUDA); } while (0)
  CUB(cudaSetDevice(device));
  ntt_init(true); stark_init();
  { const u64 w = gl_root_of_unity(NTT_TW_LOG); std::vector<u64> f = pow_series(w, 1u << (NTT_TW_LOG - 1)), b = pow_series(gl_inv(w), 1u << (NTT_TW_LOG - 1));
    CUB(cudaMalloc(&ctx->tw_fwd, f.size() * 8)); CUB(cudaMalloc(&ctx->tw_inv, b.size() * 8));
    CUB(cudaMemcpy(ctx->tw_fwd, f.data(), f.size() * 8, cudaMemcpyHostToDevice)); CUB(cudaMemcpy(ctx->tw_inv, b.data(), b.size() * 8, cudaMemcpyHostToDevice));
    CUB(cudaDeviceSynchronize()); }
  ctx->slots.resize(num_slots);
  const size_t words = slab_words_for(max_n_log2, 2, max_width), trace_words = size_t(max_width) << max_n_log2;
  for (Slot& s : ctx->slots) {
    CUB(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking)); CUB(cudaStreamCreateWithFlags(&s.copy_st, cudaStreamNonBlocking));
    for (auto& a : s.aux_st) CUB(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    for (auto& e : s.col_ev) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&s.fork_ev, cudaEventDisableTiming)); CUB(cudaEventCreateWithFlags(&s.fork2_ev, cudaEventDisableTiming)); for (auto& e : s.join_ev) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUB(cudaMalloc(&s.slab, words * 8)); s.slab_words = words;
    CUB(cudaMalloc(&s.d_state, sizeof(ProofState)));
    CUB(cudaMalloc(&s.d_partial, (size_t)(XFG_AIR_MAX_WIDTH + 2) * OOD_MAX_BLOCKS * 4 * 8)); CUB(cudaMalloc(&s.d_material, MATERIAL_WORDS * 8));
    CUB(cudaMallocHost(&s.h_state, sizeof(ProofState))); CUB(cudaMallocHost(&s.h_material, MATERIAL_WORDS * 8));
    CUB(cudaMallocHost(&s.h_seed, PROOF_INIT_BYTES)); CUB(cudaMallocHost(&s.h_trace, trace_words * 8));
    CUB(cudaMalloc(&s.d_air, sizeof(AirParams))); CUB(cudaMallocHost(&s.h_air, sizeof(AirParams)));
    CUB(cudaMalloc(&s.d_prog, sizeof(GenProgram))); CUB(cudaMallocHost(&s.h_prog, sizeof(GenProgram))); CUB(cudaMalloc(&s.d_gen, sizeof(GenState)));
    CUB(cudaMallocHost(&s.h_ood, sizeof(u64) * 2 * 2 * XFG_AIR_MAX_WIDTH));
    for (auto& e : s.ev) CUB(cudaEventCreate(&e));
  }
#undef CUB
  *out = ctx; return XFG_OK;
}

void xfg_destroy(xfg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (Slot& s : ctx->slots) {
    if (s.st) cudaStreamSynchronize(s.st);
    cudaFree(s.slab); cudaFree(s.d_state); cudaFree(s.d_partial); cudaFree(s.d_material);
    cudaFreeHost(s.h_state); cudaFreeHost(s.h_material); cudaFreeHost(s.h_seed); cudaFreeHost(s.h_trace);
    cudaFree(s.d_air); cudaFreeHost(s.h_air);
    cudaFree(s.d_prog); cudaFreeHost(s.h_prog); cudaFree(s.d_gen); cudaFreeHost(s.h_ood);
    cudaFree(s.d_go); cudaFreeHost(s.h_go); cudaFree(s.go_slab);
    for (auto& kv : s.graphs) cudaGraphExecDestroy(kv.second.exec);
    for (auto& kv : s.go_graphs) cudaGraphExecDestroy(kv.second.exec);
    for (auto& e : s.ev) if (e) cudaEventDestroy(e);
    for (auto& e : s.pev) if (e) cudaEventDestroy(e);
    for (auto& e : s.col_ev) if (e) cudaEventDestroy(e);
    if (s.fork_ev) cudaEventDestroy(s.fork_ev);
    if (s.fork2_ev) cudaEventDestroy(s.fork2_ev);
    for (auto& e : s.join_ev) if (e) cudaEventDestroy(e);
    for (auto& a : s.aux_st) if (a) cudaStreamDestroy(a);
    if (s.copy_st) cudaStreamDestroy(s.copy_st);
    if (s.st) cudaStreamDestroy(s.st);
  }
  for (auto& kv : ctx->plans) { cudaFree(kv.second.slab); cudaFree(kv.second.direct); }
  for (auto& kv : ctx->go_plans) for (void* d : kv.second.allocs) cudaFree(d);
  if (ctx->vbufs.h) cudaFreeHost(ctx->vbufs.h);
  if (ctx->vbufs.d) cudaFree(ctx->vbufs.d);
  cudaFree(ctx->tw_fwd); cudaFree(ctx->tw_inv);
  delete ctx;
}

int xfg_set_graphs(xfg_ctx* ctx, int on) { if (!ctx) return XFG_ERR_BAD_ARGS; ctx->graphs = on != 0; return XFG_OK; }
int xfg_set_profiling(xfg_ctx* ctx, int on) { if (!ctx) return XFG_ERR_BAD_ARGS; ctx->profiling = on != 0; return XFG_OK; }
int xfg_get_profile(xfg_ctx* ctx, uint32_t cap, uint32_t* count, const char** names, float* ms, uint32_t* launches) {
  if (!ctx || !count) return XFG_ERR_BAD_ARGS;
  *count = (uint32_t)ctx->prof_names.size();
  for (uint32_t i = 0; i < *count && i < cap; i++) { if (names) names[i] = ctx->prof_names[i].c_str(); if (ms) ms[i] = ctx->prof_ms[i]; if (launches) launches[i] = ctx->prof_launches[i]; }
  return XFG_OK;
}

int xfg_prove_burn_mint(xfg_ctx* ctx, const uint64_t* trace, uint32_t n_log2, const xfg_air_consts* air, const xfg_options* o, uint8_t* out,
                        size_t cap, size_t* out_len, xfg_stage_times* times) {
  return prove_common(ctx, trace, nullptr, n_log2, air, o, out, cap, out_len, times);
}
int xfg_prove_burn_mint_device(xfg_ctx* ctx, const uint64_t* d_trace, uint32_t n_log2, const xfg_air_consts* air, const xfg_options* o, uint8_t* out,
                               size_t cap, size_t* out_len, xfg_stage_times* times) {
  return prove_common(ctx, nullptr, d_trace, n_log2, air, o, out, cap, out_len, times);
}

int xfg_prove_burn_mint_cols(xfg_ctx* ctx, const uint64_t* const cols[XFG_TRACE_WIDTH], uint32_t form, uint32_t n_log2, const xfg_air_consts* air,
                             const xfg_options* o, uint8_t* out, size_t cap, size_t* out_len, xfg_stage_times* times) {
  if (!cols) return fail(ctx, XFG_ERR_BAD_ARGS, "null argument");
  return prove_common(ctx, nullptr, nullptr, n_log2, air, o, out, cap, out_len, times, false, cols, form);
}
int xfg_host_register(xfg_ctx* ctx, const void* ptr, size_t bytes) {
  if (!ctx || !ptr || !bytes) return fail(ctx, XFG_ERR_BAD_ARGS, "null argument");
  CU(cudaSetDevice(ctx->device));
  CU(cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly));
  return XFG_OK;
}
int xfg_host_unregister(xfg_ctx* ctx, const void* ptr) {
  if (!ctx || !ptr) return fail(ctx, XFG_ERR_BAD_ARGS, "null argument");
  CU(cudaSetDevice(ctx->device));
  for (Slot& s : ctx->slots) { for (auto& kv : s.graphs) cudaGraphExecDestroy(kv.second.exec); s.graphs.clear(); }   // cached graphs may hold copies from this buffer
  CU(cudaHostUnregister(const_cast<void*>(ptr)));
  return XFG_OK;
}

int xfg_prove_burn_mint_batch(xfg_ctx* ctx, uint32_t count, const uint64_t* const* traces, uint32_t n_log2, const xfg_air_consts* airs,
                              const xfg_options* o, uint8_t* out, size_t out_stride, size_t* out_lens, float* total_ms) {
  if (!ctx || !traces || !airs || !o || !out || !out_lens) return fail(ctx, XFG_ERR_BAD_ARGS, "null argument");
  if (n_log2 < MIN_LOG || n_log2 > MAX_LOG) return fail(ctx, XFG_ERR_BAD_ARGS, "trace length must be 2^3 .. 2^24");
  if (n_log2 > ctx->max_log) return fail(ctx, XFG_ERR_TOO_LARGE, "trace longer than the context was created for");
  int rc; if ((rc = check_options(ctx, o, n_log2))) return rc;
  CU(cudaSetDevice(ctx->device));
  drain_slots(ctx);
  for (uint32_t i = 0; i < count; i++) out_lens[i] = 0;   // a proof that fails (or is never started) keeps length 0
  if (go_needed(*o)) {   // general-options pipeline: one proof after the other on slot 0
    const auto t0 = std::chrono::steady_clock::now(); int first = XFG_OK; unsigned launches = 0;
    for (uint32_t i = 0; i < count; i++) {
      if ((rc = check_air(ctx, &airs[i]))) return rc;
      if (!traces[i]) return fail(ctx, XFG_ERR_BAD_ARGS, "null trace");
      rc = prove_general_burn_mint(ctx, airs[i], traces[i], nullptr, XFG_FORM_CANONICAL, nullptr, false, n_log2, *o, out + (size_t)i * out_stride, out_stride, &out_lens[i], nullptr);
      if (rc) { out_lens[i] = 0; if (!first) first = rc; }
      launches += g_xfg_launches;
    }
    g_xfg_launches = launches;
    if (total_ms) *total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return first;
  }
  const Plan* p; if ((rc = get_plan(ctx, n_log2, o->fri_remainder_max_degree, &p))) return rc;
  const int D = o->field_extension == XFG_EXT_QUADRATIC ? 2 : 1; const size_t S = ctx->slots.size();
  g_xfg_launches = 0;
  cudaEvent_t e0 = ctx->slots[0].ev[XFG_NUM_STAGES + 2], e1 = ctx->slots[0].ev[XFG_NUM_STAGES + 1];
  if (total_ms) { CU(cudaDeviceSynchronize()); CU(cudaEventRecord(e0, ctx->slots[0].st)); }
  int first_err = XFG_OK;
  for (uint32_t i = 0; i < count; i++) {
    Slot& s = ctx->slots[i % S];
    if (s.busy) { rc = finish_proof(ctx, s, out + (size_t)s.proof_index * out_stride, out_stride, &out_lens[s.proof_index], nullptr); if (rc && !first_err) first_err = rc; }
    if ((rc = check_air(ctx, &airs[i]))) { drain_slots(ctx); return rc; }
    if (!traces[i]) { drain_slots(ctx); return fail(ctx, XFG_ERR_BAD_ARGS, "null trace"); }
    s.generic = false; s.W = XFG_TRACE_WIDTH; s.fill_trace = false; s.in_scale = 1; s.tail_threads = 256;
    if ((rc = upload_trace(ctx, s, *p, D, traces[i], false))) { drain_slots(ctx); return rc; }
    s.proof_index = i;
    if ((rc = launch_proof(ctx, s, *p, D, *o, airs[i], nullptr, false))) { drain_slots(ctx); return rc; }
  }
  for (Slot& s : ctx->slots) if (s.busy) { rc = finish_proof(ctx, s, out + (size_t)s.proof_index * out_stride, out_stride, &out_lens[s.proof_index], nullptr); if (rc && !first_err) first_err = rc; }
  if (total_ms) { CU(cudaDeviceSynchronize()); CU(cudaEventRecord(e1, ctx->slots[0].st)); CU(cudaEventSynchronize(e1)); cudaEventElapsedTime(total_ms, e0, e1); }
  return first_err;
}

int xfg_burn_mint_pack_inputs(xfg_ctx* ctx, uint64_t burn, uint64_t mint, const uint8_t txp[32], const uint8_t* rcpt, size_t rcpt_len, const uint8_t* secret,
                              size_t secret_len, uint32_t network_id, uint32_t target_chain_id, uint32_t version, xfg_air_consts* out) {
  if (!txp || !rcpt || !secret || !out) return fail(ctx, XFG_ERR_BAD_ARGS, "null argument");
  std::string err; int rc = burn_mint_pack_inputs(burn, mint, txp, rcpt, rcpt_len, secret, secret_len, network_id, target_chain_id, version, out, err);
  if (rc && ctx) ctx->last_error = err;
  return rc;
}
int xfg_burn_mint_build_trace(const xfg_air_consts* air, uint32_t n_log2, uint64_t* t) {
  if (!air || !t || n_log2 < MIN_LOG || n_log2 > MAX_LOG) return XFG_ERR_BAD_ARGS;
  burn_mint_build_trace(air, n_log2, t); return XFG_OK;
}
int xfg_prove_burn_mint_from_inputs(xfg_ctx* ctx, uint64_t burn, uint64_t mint, const uint8_t txp[32], const uint8_t* rcpt, size_t rcpt_len,
                                    const uint8_t* secret, size_t secret_len, uint32_t network_id, uint32_t target_chain_id, uint32_t version,
                                    uint32_t n_log2, const xfg_options* o, uint8_t* out, size_t cap, size_t* out_len, xfg_stage_times* times) {
  if (!ctx) return XFG_ERR_BAD_ARGS;
  if (n_log2 < MIN_LOG || n_log2 > ctx->max_log) return fail(ctx, XFG_ERR_TOO_LARGE, "trace longer than the context was created for");
  xfg_air_consts air; int rc = xfg_burn_mint_pack_inputs(ctx, burn, mint, txp, rcpt, rcpt_len, secret, secret_len, network_id, target_chain_id, version, &air);
  if (rc) return rc;
  // build_trace (src/burn_mint_air.rs:442-476) runs on the device: six constant columns and the state step function need no upload
  return prove_common(ctx, nullptr, nullptr, n_log2, &air, o, out, cap, out_len, times, true);
}

}  // extern "C"

#include "stage_api.inc"
#include "wide.inc"
#include "verify_api.inc"
#include "air_api.inc"
#include "general_api.inc"
