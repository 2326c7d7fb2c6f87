// general_pipeline.cuh — plan, workspace layout and launch sequence of the general-options proof (see general_bodies.cuh), written against
// a backend `BK` that knows how to run a body over an index space (`run`), a batched NTT job (`ntt`) and the Merkle levels above a leaf
// level (`merkle_upper`).  The product's backend launches CUDA kernels on the proof's stream (general_api.inc); tests/host_emul supplies a
// plain-C++ backend so that the very same sequence and bodies are checked on a CPU-only box.
//
// Stage order = winter-prover 0.8.3 `Prover::prove` / `generate_proof`, as in prover.cu: enqueue_proof.
#pragma once
#include <utility>
#include <vector>
#include "general_bodies.cuh"
#include "proof_bytes.hpp"

namespace xfg {

struct GoPlan {
  u32 ln = 0, lb = 0, lf = 0, lN = 0, lce = 1; size_t n = 0, N = 0;      // lce: log2 of the constraint-evaluation blowup (1 for degrees <= 3, 2 for 4-5, 3 for 6-9)
  u32 num_layers = 0, layer_log[GO_MAX_LAYERS + 1] = {0}, rem_log = 0, rem_len = 0;
  NttTables ntt{};                                     // tw_* (context-wide), wn_fwd / wn_inv
  PowTable wN_inv{};
  const u64 *pre_lo = nullptr, *pre_hi = nullptr; u32 pre_hi_stride = 0;   // s_k = 7 w_N^k, k < B
  const u64 *un_lo = nullptr, *un_hi = nullptr; u32 un_hi_stride = 0;      // (7 w_(ce n)^c)^-1, c < ce: the cosets of the constraint-evaluation domain
  const u64* d_sk = nullptr;
  u64 s_ce[GO_MAX_CE] = {0}, zinv[GO_MAX_CE] = {0}, g_n = 0, g_last = 0, n_inv = 0, nr_inv = 0;
  GoFriConsts fc{}; GoCombineConsts cc{};
};

// FriOptions::num_fri_layers (A.10) and the shapes the reference itself refuses with a panic (tests/golden/reference_proofs_options.json: "refused")
// returns nullptr when the shape is fine, else the reason
// max_degree: highest transition-constraint degree of the AIR; ce_blowup = max(2, next_pow2(max_degree - 1)) (TransitionConstraintDegree::min_blowup_factor, A.3)
inline u32 go_ce_log(u32 max_degree) { u32 l = 1; while ((1u << l) + 1 < max_degree) l++; return l; }
inline u32 go_comp_columns(u32 max_degree) { return max_degree > 2 ? max_degree - 1 : 1; }     // AirContext::num_constraint_composition_columns
inline const char* go_plan_shape(GoPlan& p, u32 ln, u32 blowup, u32 folding, u32 rem_max_deg, u32 max_degree = 2) {
  auto lg = [](u32 x) { u32 r = 0; while ((1u << r) < x) r++; return r; };
  p.ln = ln; p.lb = lg(blowup); p.lf = lg(folding); p.lN = ln + p.lb; p.n = size_t(1) << ln; p.N = size_t(1) << p.lN; p.lce = go_ce_log(max_degree);
  if (p.lce > p.lb) return "blowup factor too small for the degree of the transition constraints";
  const size_t mx = (size_t)(rem_max_deg + 1) * blowup;
  u32 l = p.lN; p.num_layers = 0; p.layer_log[0] = l;
  while ((size_t(1) << l) > mx) {
    if (l < p.lf + 1) return "failed to construct FRI layer tree: a layer of fewer than two rows (TooFewLeaves)";
    if (p.num_layers == GO_MAX_LAYERS) return "too many FRI layers";
    l -= p.lf; p.layer_log[++p.num_layers] = l;
  }
  p.rem_log = l;
  if (l < p.lb) return "the FRI remainder polynomial would be empty (remainder domain smaller than the blowup factor)";
  p.rem_len = 1u << (l - p.lb);
  if (p.rem_len > (u32)MAX_REMAINDER) return "FRI remainder too long";
  return nullptr;
}

// host-side values of the plan's tables; `up(vector)` places one table where the backend's bodies can read it and returns the pointer
template <class Up> inline void go_plan_tables(GoPlan& p, Up up) {
  auto series = [](u64 base, size_t count) { std::vector<u64> v(count); u64 x = 1; for (size_t i = 0; i < count; i++) { v[i] = x; x = gl_mul(x, base); } return v; };
  const u32 B = 1u << p.lb, F = 1u << p.lf;
  const u32 nhi_n = (u32)std::max<size_t>(1, p.n >> POW_LO_BITS), nhi_N = (u32)std::max<size_t>(1, p.N >> POW_LO_BITS);
  const u64 wN = gl_root_of_unity(p.lN), wn = gl_root_of_unity(p.ln);
  auto table = [&](u64 base, u32 nhi) { PowTable t; t.lo = up(series(base, POW_LO)); t.hi = up(series(gl_pow(base, POW_LO), nhi)); return t; };
  p.ntt.wn_fwd = table(wn, nhi_n); p.ntt.wn_inv = table(gl_inv(wn), nhi_n); p.wN_inv = table(gl_inv(wN), nhi_N);
  std::vector<u64> sk(B), lo, hi;
  for (u32 k = 0; k < B; k++) {
    sk[k] = gl_mul(XFG_GENERATOR, gl_pow(wN, k));
    const std::vector<u64> a = series(sk[k], POW_LO), b = series(gl_pow(sk[k], POW_LO), nhi_n);
    lo.insert(lo.end(), a.begin(), a.end()); hi.insert(hi.end(), b.begin(), b.end());
  }
  p.pre_lo = up(lo); p.pre_hi = up(hi); p.pre_hi_stride = nhi_n; p.d_sk = up(sk);
  const u32 ce = 1u << p.lce;
  lo.clear(); hi.clear();
  for (u32 c = 0; c < ce; c++) {
    p.s_ce[c] = sk[(size_t)c * (B / ce)];
    const u64 inv = gl_inv(p.s_ce[c]);
    const std::vector<u64> a = series(inv, POW_LO), b = series(gl_pow(inv, POW_LO), nhi_n);
    lo.insert(lo.end(), a.begin(), a.end()); hi.insert(hi.end(), b.begin(), b.end());
    p.zinv[c] = gl_inv(gl_sub(gl_pow(p.s_ce[c], p.n), 1));
  }
  p.un_lo = up(lo); p.un_hi = up(hi); p.un_hi_stride = nhi_n;
  p.g_n = wn; p.g_last = gl_pow(wn, p.n - 1); p.n_inv = gl_inv((u64)p.n); p.nr_inv = gl_inv(u64(1) << p.rem_log);
  { const u64 wci = gl_inv(gl_root_of_unity(p.lce)), g7ni = gl_inv(gl_pow(XFG_GENERATOR, p.n)), cei = gl_inv(ce);
    for (u32 e = 0; e < (u32)GO_MAX_CE; e++) { p.cc.wi[e] = e < ce ? gl_pow(wci, e) : 0; p.cc.scale[e] = e < ce ? gl_mul(cei, gl_pow(g7ni, e)) : 0; } }
  const u64 wfi = gl_inv(gl_root_of_unity(p.lf));
  for (u32 j = 0; j < 16; j++) p.fc.wfi[j] = j < F ? gl_pow(wfi, j) : 0;
  p.fc.f_inv = gl_inv(F); p.fc.inv7 = gl_inv(XFG_GENERATOR);
}

// workspace of one proof (all offsets in u64 words from the base; every region 64-byte aligned)
struct GoCarve {
  u64 *trace_in, *trace_coef, *lde, *ce, *ce_tmp, *h_coef, *h_lde, *deep, *ood_partial, *ood_part2, *ood_sums;
  Digest *trace_tree, *comp_tree;
  u64* fri_evals[GO_MAX_LAYERS + 1]; Digest* fri_tree[GO_MAX_LAYERS];
  size_t words;
};
// gap / gaps: optional guard words after every region (the emulation harness poisons them under AddressSanitizer: tests/host_emul)
inline void go_carve(u64* base, const GoPlan& p, int D, u32 W, u32 K, GoCarve& c, size_t gap = 0, std::vector<std::pair<size_t, size_t>>* gaps = nullptr) {
  u64* w = base; const size_t n = p.n, N = p.N;
  auto take = [&](size_t k) { u64* r = w; w += (k + 7) & ~size_t(7); if (gaps) gaps->push_back({(size_t)(w - base), gap}); w += gap; return r; };
  c.trace_in = take(W * n); c.trace_coef = take(W * n); c.lde = take(W * N);
  c.trace_tree = reinterpret_cast<Digest*>(take(8 * N));
  const size_t ce = size_t(1) << p.lce;
  c.ce = take(ce * D * n); c.ce_tmp = take(ce * D * n); c.h_coef = take((size_t)K * D * n); c.h_lde = take((size_t)K * D * N);
  c.comp_tree = reinterpret_cast<Digest*>(take(8 * N));
  c.deep = take(D * N); c.fri_evals[0] = c.deep;
  for (u32 l = 1; l <= p.num_layers; l++) c.fri_evals[l] = take((size_t)D << p.layer_log[l]);
  for (u32 l = 0; l < p.num_layers; l++) c.fri_tree[l] = reinterpret_cast<Digest*>(take((size_t)8 << (p.layer_log[l] - p.lf)));   // 2 * Nl/F digests
  const size_t P = W + (size_t)K * D;      // polynomials of the out-of-domain evaluation: trace columns, then the limbs of the composition columns
  c.ood_partial = take(P * std::min<size_t>(GO_OOD_CHUNKS, n) * 2 * GO_MAX_EXT); c.ood_part2 = take(P * 2 * 32 * GO_MAX_EXT); c.ood_sums = take(P * 2 * GO_MAX_EXT);
  c.words = (size_t)(w - base);
}

// layout of the material buffer (opened rows + per-position sibling paths); returns its size in words
inline size_t go_gather_tasks(const GoPlan& p, int D, u32 W, u32 K, u32 q, const GoCarve& c, std::vector<GoGatherTask>& tasks) {
  size_t off = 0; tasks.clear();
  auto add = [&](const u64* src, const Digest* tree, u64 limb_stride, u32 coset, u64 R, u64 M, u32 J, u32 limbs, u32 depth, int layer) {
    GoGatherTask k{}; k.src = src; k.tree = tree; k.limb_stride = limb_stride; k.coset = coset; k.lb = p.lb; k.ln = p.ln; k.R = R; k.M = M; k.J = J; k.limbs = limbs;
    k.depth = depth; k.fri_layer = layer; k.max_q = q; k.rows_off = off; off += ((size_t)q * J * limbs + 3) & ~size_t(3); k.paths_off = off; off += (size_t)q * depth * 4;
    tasks.push_back(k);
  };
  add(c.lde, c.trace_tree, p.N, 1, 0, p.N, 1, W, p.lN, -1);
  add(c.h_lde, c.comp_tree, p.N, 1, 0, p.N, 1, K * (u32)D, p.lN, -1);
  for (u32 l = 0; l < p.num_layers; l++) {
    const u64 Nl = u64(1) << p.layer_log[l], R = Nl >> p.lf;
    add(c.fri_evals[l], c.fri_tree[l], Nl, 0, R, R, 1u << p.lf, (u32)D, p.layer_log[l] - p.lf, (int)l);
  }
  return off;
}

// the whole proof, enqueued on the backend (no host synchronisation); the trace is at `trace_src` (c.trace_in or a caller's device buffer)
template <int D, class BK>
void go_enqueue_d(BK& bk, const GoPlan& p, const GoCarve& c, GoState* s, const GenProgram* prog, u32 W, u32 K, u32 num_assertions, u32 ncoef, const u64* trace_src, u64 in_scale,
                  u32 num_queries, u32 grinding, const std::vector<GoGatherTask>& tasks, u64* material) {
  const u32 ln = p.ln, B = 1u << p.lb, ce = 1u << p.lce; const size_t n = p.n, N = p.N;
  // (where the composition columns leave no vanishing coefficient to check - K = ce - the trace itself is validated first)
  if (K == ce && in_scale == 1) bk.run(n - 1 + num_assertions, GoValidate{trace_src, ln, prog, s});
  // 1 ---- extend_execution_trace: interpolate, evaluate on the B cosets s_k <w_n>; every trace element must be canonical (checked by the first pass)
  { NttJob j{}; j.src = trace_src; j.dst = c.trace_coef; j.ln = ln; j.batch = W; j.src_tstride = n; j.dst_tstride = n; j.src_div = 1;
    j.canon_flag = &s->error_flags; j.canon_bit = ERR_FLAG_NONCANONICAL; j.inverse = true; j.scale = gl_mul(p.n_inv, in_scale); bk.ntt(j); }
  { NttJob j{}; j.src = c.trace_coef; j.dst = c.lde; j.ln = ln; j.batch = W * B; j.src_tstride = n; j.dst_tstride = n; j.src_div = B;
    j.inverse = false; j.scale = 1; j.pre_lo = p.pre_lo; j.pre_hi = p.pre_hi; j.pre_hi_stride = p.pre_hi_stride; bk.ntt(j); }
  //   ---- compute_execution_trace_commitment, commit_trace, composition coefficients
  bk.run(N, GoLeaf{c.lde, N, W, p.lb, ln, c.trace_tree});
  bk.merkle_upper(c.trace_tree, N);
  bk.run(1, GoStepTrace<D>{s, c.trace_tree, ncoef});
  // 2 ---- evaluate_constraints
  const u32 pts = n >= 1024 ? GO_PTS : 1;   // points per thread of the constraint / DEEP bodies (batched inversions); short traces keep one point per thread
  { GoConstraint<D> k{}; k.lde = c.lde; k.ln = ln; k.lb = p.lb; k.lce = p.lce; k.pts = pts; k.prog = prog; k.s = s; k.wn = p.ntt.wn_fwd;
    for (u32 q = 0; q < ce; q++) { k.s_ce[q] = p.s_ce[q]; k.zinv[q] = p.zinv[q]; }
    k.g_last = p.g_last; k.out = c.ce; bk.run((size_t)ce * n / pts, k); }
  // 3 ---- commit_to_constraint_evaluations: coset interpolation (ce cosets of n points), K composition columns, LDE, commitment
  { NttJob j{}; j.src = c.ce; j.dst = c.ce_tmp; j.ln = ln; j.batch = ce * D; j.src_tstride = n; j.dst_tstride = n; j.src_div = 1;
    j.inverse = true; j.scale = p.n_inv; j.post_lo = p.un_lo; j.post_hi = p.un_hi; j.post_hi_stride = p.un_hi_stride; j.post_div = ce; bk.ntt(j); }
  bk.run(n, GoCombine{c.ce_tmp, ln, p.lce, K, D, p.cc, c.h_coef, s});
  { NttJob j{}; j.src = c.h_coef; j.dst = c.h_lde; j.ln = ln; j.batch = K * D * B; j.src_tstride = n; j.dst_tstride = n; j.src_div = B;
    j.inverse = false; j.scale = 1; j.pre_lo = p.pre_lo; j.pre_hi = p.pre_hi; j.pre_hi_stride = p.pre_hi_stride; bk.ntt(j); }
  bk.run(N, GoLeaf{c.h_lde, N, K * (u32)D, p.lb, ln, c.comp_tree});
  bk.merkle_upper(c.comp_tree, N);
  bk.run(1, GoStepComp<D>{s, c.comp_tree, p.g_n});
  // 4 ---- build_deep_composition_poly: OOD frame + DEEP coefficients
  const u32 chunks = (u32)std::min<size_t>(GO_OOD_CHUNKS, n);
  const size_t P = W + (size_t)K * D;
  bk.run(P * chunks, GoOodPartial<D>{c.trace_coef, c.h_coef, ln, W, chunks, s, c.ood_partial});
  { const u32 groups = chunks >= 64 ? 32 : 1;      // two-level sum: 32 group sums per (polynomial, point), then their sum
    if (groups > 1) { bk.run(P * 2 * groups, GoOodSum<D>{c.ood_partial, chunks, groups, c.ood_part2}); bk.run(P * 2, GoOodSum2<D>{c.ood_part2, groups, c.ood_sums}); }
    else bk.run(P * 2, GoOodSum<D>{c.ood_partial, chunks, 1, c.ood_sums}); }
  bk.run(1, GoStepOod<D>{s, c.ood_sums, W, K});
  // 5 ---- evaluate_deep_composition_poly (pointwise)
  bk.run(N / pts, GoDeep<D>{c.lde, c.h_lde, ln, p.lb, W, K, pts, s, p.ntt.wn_fwd, p.d_sk, c.deep});
  // 6 ---- compute_fri_layers
  for (u32 l = 0; l < p.num_layers; l++) {
    const u32 lNl = p.layer_log[l]; const u64 Nl = u64(1) << lNl, R = Nl >> p.lf;
    bk.run(R, GoFriLeaf<D>{c.fri_evals[l], Nl, 1u << p.lf, R, c.fri_tree[l]});
    bk.merkle_upper(c.fri_tree[l], R);
    bk.run(1, GoStepFri<D>{s, c.fri_tree[l], l});
    bk.run(R, GoFriFold<D>{c.fri_evals[l], Nl, lNl, p.lf, l, s, p.wN_inv, p.lN, p.fc, c.fri_evals[l + 1], R});
  }
  bk.run((size_t)p.rem_len * D, GoRemainder<D>{c.fri_evals[p.num_layers], u64(1) << p.rem_log, p.rem_log, p.lN, p.wN_inv, p.nr_inv, p.fc.inv7, s});
  bk.run(1, GoStepRemainder<D>{s, p.rem_len});
  // 7 ---- determine_query_positions
  { const u64 total = bk.grind_threads(grinding); bk.run(total, GoGrind{s, grinding, total}); }
  bk.run(1, GoStepPositions{s, num_queries, p.lN});
  if (p.num_layers) bk.run(p.num_layers, GoFoldPositions{s, p.lN, p.lf});
  // 8 ---- build_proof_object: opened rows + authentication nodes
  for (const GoGatherTask& k : tasks) bk.run((size_t)k.max_q * (k.J * k.limbs + k.depth), GoGather{k, s, material});
}
template <class BK>
void go_enqueue(BK& bk, int D, const GoPlan& p, const GoCarve& c, GoState* s, const GenProgram* prog, u32 W, u32 K, u32 num_assertions, u32 ncoef, const u64* trace_src, u64 in_scale,
                u32 num_queries, u32 grinding, const std::vector<GoGatherTask>& tasks, u64* material) {
  if (D == 1) go_enqueue_d<1>(bk, p, c, s, prog, W, K, num_assertions, ncoef, trace_src, in_scale, num_queries, grinding, tasks, material);
  else if (D == 2) go_enqueue_d<2>(bk, p, c, s, prog, W, K, num_assertions, ncoef, trace_src, in_scale, num_queries, grinding, tasks, material);
  else go_enqueue_d<3>(bk, p, c, s, prog, W, K, num_assertions, ncoef, trace_src, in_scale, num_queries, grinding, tasks, material);
}

// StarkProof::to_bytes (A.12) from the proof state and the gathered material (host side)
inline void go_assemble(const GoPlan& p, int D, u32 W, u32 K, const xfg_options& o, const GoState& s, const u64* mat, const std::vector<GoGatherTask>& g, std::vector<u8>& bytes) {
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
  // Queries: u32 len + values, u32 len + paths
  auto queries = [&](const GoGatherTask& t, const u32* pos, u32 cnt) {
    const size_t vbytes = (size_t)cnt * t.J * t.limbs * 8;
    out.u32_(vbytes); out.raw(mat + t.rows_off, vbytes);
    Out pth; batch_paths(pos, cnt, mat + t.paths_off, t.depth, t.M, pth);
    out.u32_(pth.b.size()); out.raw(pth.b.data(), pth.b.size());
  };
  queries(g[0], s.positions, s.num_positions);
  queries(g[1], s.positions, s.num_positions);
  // OodFrame
  out.u16_(1 + 2 * (size_t)W * D * 8); out.u8_(2);
  for (u32 i = 0; i < 2 * W; i++) for (int l = 0; l < D; l++) out.u64_(s.ood_frame[i][l]);
  out.u16_((size_t)K * D * 8); for (u32 i = 0; i < K; i++) for (int l = 0; l < D; l++) out.u64_(s.hz[i][l]);
  // FriProof
  out.u8_(p.num_layers);
  for (u32 l = 0; l < p.num_layers; l++) queries(g[2 + l], s.fri_positions[l], s.fri_num_positions[l]);
  out.u16_((size_t)s.remainder_len * D * 8);
  for (u32 i = 0; i < s.remainder_len; i++) for (int l = 0; l < D; l++) out.u64_(s.remainder[i][l]);
  out.u8_(0);   // FriProof::num_partitions as log2 (one partition), pinned against the reference binary
  out.u64_(s.nonce);
  bytes.swap(out.b);
}

}  // namespace xfg
