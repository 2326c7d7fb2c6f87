// oracle/prover.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// CPU restatement of winter-prover 0.8.3 `Prover::prove` / `generate_proof` for the normalised BurnMintAir
// (call site: src/burn_mint_prover.rs:124-126; hooks src/burn_mint_air.rs:479-531).  Stage order and every
// transcript / serialisation detail follow SURVEY.md §3.1 and Appendix A (A.4-A.12).  Deliberately structured
// like Winterfell (coefficient-domain DEEP with synthetic division, whole-vector FRI folding) so that the GPU
// product - which computes the same values pointwise - is checked by an independently shaped computation.
//
// PARITY STATUS: PINNED.  The upstream crates are not in /root/reference, but the reference's shipped arm64 binary
// (test-dist/xfg-stark-cli, winterfell 0.8.3 linked in) is executed here by oracle/a64emu; the proofs its prover emits
// (tests/golden/reference_proofs.json) are byte-equal to this file's output on the same statement, trace and options
// (tests/test_reference_binary_pins.py).  Also pinned: BLAKE3 (Python blake3), Keccak (src/lib.rs:141-148 KAT), field/NTT algebra (big-int).
#pragma once
#include <algorithm>
#include <chrono>
#include "air.hpp"
#include "coin.hpp"
#include "merkle.hpp"
#include "ntt.hpp"

namespace orc {

extern int g_threads;   // OpenMP threads for the data-parallel loops (1 = the reference's actual serial build)

enum Stage { ST_EXTEND_TRACE = 0, ST_COMMIT_TRACE, ST_EVAL_CONSTRAINTS, ST_COMMIT_CONSTRAINTS, ST_BUILD_DEEP, ST_EVAL_DEEP,
             ST_FRI_LAYERS, ST_QUERY_POSITIONS, ST_BUILD_PROOF, ST_COUNT };
static const char* const STAGE_NAMES[ST_COUNT] = {   // winter-prover tracing span names
  "extend_execution_trace", "compute_execution_trace_commitment", "evaluate_constraints", "commit_to_constraint_evaluations",
  "build_deep_composition_poly", "evaluate_deep_composition_poly", "compute_fri_layers", "determine_query_positions", "build_proof_object"};
struct StageTimes { double ms[ST_COUNT] = {0}; };
struct Timer { std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double lap() { auto t1 = std::chrono::steady_clock::now(); double r = std::chrono::duration<double, std::milli>(t1 - t0).count(); t0 = t1; return r; } };

// intermediate values exposed for stage-level parity tests
template <class E> struct ProverDebug {
  Digest trace_root{}, constraint_root{}, remainder_commitment{};
  std::vector<Digest> fri_roots;
  std::vector<E> tcoef, bcoef, dcoef, alphas, ood_frame, remainder, ce_evals, deep_evals;
  E z{}, hz{};
  u64 nonce = 0;
  std::vector<size_t> positions;
};

inline void put_u16(std::vector<u8>& o, size_t v) { o.push_back((u8)v); o.push_back((u8)(v >> 8)); }
inline void put_u32(std::vector<u8>& o, size_t v) { for (int i = 0; i < 4; i++) o.push_back((u8)(v >> (8 * i))); }
inline void put_bytes(std::vector<u8>& o, const std::vector<u8>& b) { o.insert(o.end(), b.begin(), b.end()); }

// winter-fri folding::fold_positions (A.10, D): order-preserving dedup, not re-sorted
inline std::vector<size_t> fold_positions(const std::vector<size_t>& pos, size_t source_domain, size_t folding) {
  size_t target = source_domain / folding; std::vector<size_t> r;
  for (size_t p : pos) { size_t q = p % target; if (std::find(r.begin(), r.end(), q) == r.end()) r.push_back(q); }
  return r;
}

// winter-fri folding::apply_drp on one row: interpolate over x * w_F^j and evaluate at alpha (A.10)
template <class E> inline E fold_row(const E* row, size_t F, u64 x_inv, E alpha) {
  std::vector<E> p(row, row + F);
  ntt_core(p.data(), F, true);
  u64 off = finv((u64)F); E r = E::zero();
  std::vector<E> c(F);
  for (size_t k = 0; k < F; k++) { c[k] = p[k].mul_base(off); off = fmul(off, x_inv); }
  for (size_t k = F; k-- > 0;) r = r * alpha + c[k];
  return r;
}

template <class E> struct FriLayerData { std::vector<E> evals; MerkleTree tree; };

template <class E>
std::vector<u8> prove(const std::vector<std::vector<F1>>& trace, const AirDef& air,
                      const ProofOptions& opt, StageTimes* times = nullptr, ProverDebug<E>* dbg = nullptr) {
  const size_t W = air.width, NT = air.num_transition, n = trace[0].size(), b = opt.blowup, N = n * b, c = air.ce_blowup(), K = air.comp_columns(), F = opt.folding;
  if (trace.size() != W) throw std::runtime_error("trace width does not match the AIR");
  if (n < 8 || (n & (n - 1))) throw std::runtime_error("trace length must be a power of two >= 8");
  if (b < c) throw std::runtime_error("blowup factor too small");
  const int T = g_threads; (void)T;
  const u64 g_n = root_of_unity(ilog2(n)), g_ce = root_of_unity(ilog2(c * n));
  const u64 offset = XFG_GENERATOR;
  twiddles(n, false); twiddles(n, true); twiddles(c * n, true); twiddles(F, true);   // fill caches before threading
  StageTimes st; Timer tm;

  // 0 ----- channel: coin seeded with context + public inputs (A.4)
  RandomCoin coin(seed_elements(n, opt, W, air.pub_inputs));

  // 1 ----- extend_execution_trace: interpolate columns, evaluate over the LDE coset (A.7)
  std::vector<std::vector<F1>> polys(trace), lde(W);
#pragma omp parallel for num_threads(T) schedule(dynamic)
  for (size_t j = 0; j < W; j++) { interpolate_poly(polys[j]); lde[j] = evaluate_poly_with_offset(polys[j], offset, b); }
  st.ms[ST_EXTEND_TRACE] = tm.lap();
  //       compute_execution_trace_commitment: leaf = hash of the 7 real elements of the row (A.7, D)
  std::vector<Digest> leaves(N);
#pragma omp parallel for num_threads(T) schedule(static)
  for (size_t i = 0; i < N; i++) { std::vector<F1> row(W); for (size_t j = 0; j < W; j++) row[j] = lde[j][i]; leaves[i] = hash_elements(row.data(), W); }
  MerkleTree trace_tree(std::move(leaves));
  coin.reseed(trace_tree.root());
  st.ms[ST_COMMIT_TRACE] = tm.lap();

  // 2 ----- evaluate_constraints over the constraint-evaluation domain (A.8)
  const std::vector<Assertion>& asr = air.assertions;
  std::vector<E> tcoef(NT), bcoef(asr.size());
  for (auto& x : tcoef) x = coin.draw<E>();
  for (auto& x : bcoef) x = coin.draw<E>();
  // boundary constraint groups: one per distinct divisor (x - g^step) (A.8)
  std::vector<size_t> gsteps, gof(asr.size());
  for (size_t k = 0; k < asr.size(); k++) { size_t g = std::find(gsteps.begin(), gsteps.end(), asr[k].step) - gsteps.begin(); if (g == gsteps.size()) gsteps.push_back(asr[k].step); gof[k] = g; }
  const size_t G = gsteps.size();
  const size_t cn = c * n, lde_shift = b / c;
  const u64 g_last = fpow(g_n, n - 1);                 // exemption point g^(n-1)
  std::vector<u64> xs = power_series(g_ce, cn, offset);
  // divisor inverses: (x^n - 1) is periodic with period c; (x - 1) and (x - g^(n-1)) need a batch inversion
  std::vector<F1> zt(c); std::vector<std::vector<F1>> dg(G, std::vector<F1>(cn));
  for (size_t s = 0; s < c; s++) zt[s] = F1(fsub(fpow(xs[s], n), 1));
  for (size_t g = 0; g < G; g++) { const u64 pt = fpow(g_n, gsteps[g]); for (size_t s = 0; s < cn; s++) dg[g][s] = F1(fsub(xs[s], pt)); dg[g] = batch_inverse(dg[g]); }
  zt = batch_inverse(zt);
  std::vector<E> hev(cn);
#pragma omp parallel for num_threads(T) schedule(static)
  for (size_t s = 0; s < cn; s++) {
    size_t i0 = s * lde_shift, i1 = (i0 + b) % N;      // frame rows (A.8, D)
    std::vector<F1> cur(W), nxt(W), r(NT);
    for (size_t j = 0; j < W; j++) { cur[j] = lde[j][i0]; nxt[j] = lde[j][i1]; }
    eval_air_transition<F1>(air, cur.data(), nxt.data(), r.data());
    E tsum = E::zero(); std::vector<E> bs(G, E::zero());
    for (size_t k = 0; k < NT; k++) tsum = tsum + tcoef[k].mul_base(r[k].v);
    for (size_t k = 0; k < asr.size(); k++) bs[gof[k]] = bs[gof[k]] + bcoef[k].mul_base(fsub(cur[asr[k].column].v, asr[k].value));
    E h = tsum.mul_base(fmul(fsub(xs[s], g_last), zt[s % c].v));
    for (size_t g = 0; g < G; g++) h = h + bs[g].mul_base(dg[g][s].v);
    hev[s] = h;
  }
  if (dbg) dbg->ce_evals = hev;
  st.ms[ST_EVAL_CONSTRAINTS] = tm.lap();

  // 3 ----- commit_to_constraint_evaluations: composition poly, its LDE and commitment (A.9)
  interpolate_poly_with_offset(hev, offset);
  for (size_t k = K * n; k < cn; k++) if (!hev[k].is_zero()) throw std::runtime_error("UnsatisfiedTransitionConstraintError");
  // CompositionPoly::new: column i = coefficients [i n, (i + 1) n) of the composition polynomial (K = 1 for degrees <= 2)
  std::vector<std::vector<E>> hcols(K), hldes(K);
  for (size_t i = 0; i < K; i++) { hcols[i].assign(hev.begin() + i * n, hev.begin() + (i + 1) * n); hldes[i] = evaluate_poly_with_offset(hcols[i], offset, b); }
  std::vector<Digest> cleaves(N);
#pragma omp parallel for num_threads(T) schedule(static)
  for (size_t i = 0; i < N; i++) { std::vector<E> row(K); for (size_t j = 0; j < K; j++) row[j] = hldes[j][i]; cleaves[i] = hash_elements(row.data(), K); }
  MerkleTree ctree(std::move(cleaves));
  coin.reseed(ctree.root());
  st.ms[ST_COMMIT_CONSTRAINTS] = tm.lap();

  // 4 ----- build_deep_composition_poly (A.9): OOD frame, DEEP coefficients, quotients in coefficient form
  E z = coin.draw<E>(), zg = z.mul_base(g_n);
  std::vector<E> tz(W), tzg(W), frame;
#pragma omp parallel for num_threads(T) schedule(dynamic)
  for (size_t j = 0; j < W; j++) { tz[j] = eval_poly<E, F1>(polys[j], z); tzg[j] = eval_poly<E, F1>(polys[j], zg); }
  for (size_t j = 0; j < W; j++) { frame.push_back(tz[j]); frame.push_back(tzg[j]); }   // interleaved per column (D)
  std::vector<u8> ood_trace_bytes; ood_trace_bytes.push_back(2); put_elems(ood_trace_bytes, frame);
  coin.reseed(hash_elements(frame));
  std::vector<E> hzs(K);
  for (size_t i = 0; i < K; i++) hzs[i] = eval_poly<E, E>(hcols[i], z);
  const E hz = hzs[0];
  std::vector<u8> ood_eval_bytes; put_elems(ood_eval_bytes, hzs);
  coin.reseed(hash_elements(hzs));
  std::vector<E> dcoef(W + K);
  for (auto& x : dcoef) x = coin.draw<E>();
  std::vector<E> t1(n), t2(n);
#pragma omp parallel for num_threads(T) schedule(static)
  for (size_t k = 0; k < n; k++) { E a = E::zero(); for (size_t j = 0; j < W; j++) a = a + dcoef[j].mul_base(polys[j][k].v); t1[k] = a; }
  t2 = t1;
  { E s1 = E::zero(), s2 = E::zero(); for (size_t j = 0; j < W; j++) { s1 = s1 + dcoef[j] * tz[j]; s2 = s2 + dcoef[j] * tzg[j]; }
    t1[0] = t1[0] - s1; t2[0] = t2[0] - s2; }
  syn_div_in_place(t1, z); syn_div_in_place(t2, zg);
  std::vector<E> deep(n);
  for (size_t k = 0; k < n; k++) deep[k] = t1[k] + t2[k];
  for (size_t i = 0; i < K; i++) {
    std::vector<E> hq(hcols[i]); hq[0] = hq[0] - hzs[i]; syn_div_in_place(hq, z);
    for (size_t k = 0; k < n; k++) deep[k] = deep[k] + dcoef[W + i] * hq[k];
  }
  st.ms[ST_BUILD_DEEP] = tm.lap();

  // 5 ----- evaluate_deep_composition_poly over the LDE domain
  std::vector<E> evals = evaluate_poly_with_offset(deep, offset, b);
  if (dbg) dbg->deep_evals = evals;
  st.ms[ST_EVAL_DEEP] = tm.lap();

  // 6 ----- compute_fri_layers (A.10): constant domain offset 7 at every layer
  std::vector<FriLayerData<E>> layers; std::vector<E> alphas; std::vector<Digest> fri_roots;
  const size_t num_layers = opt.num_fri_layers(N);
  for (size_t l = 0; l < num_layers; l++) {
    size_t Nl = evals.size(), rows = Nl / F;
    std::vector<E> tr(Nl);                               // transpose_slice: row i = [v_i, v_{i+rows}, ...]
    std::vector<Digest> lv(rows);
#pragma omp parallel for num_threads(T) schedule(static)
    for (size_t i = 0; i < rows; i++) { for (size_t j = 0; j < F; j++) tr[i * F + j] = evals[i + j * rows]; lv[i] = hash_elements(&tr[i * F], F); }
    MerkleTree tree(std::move(lv));
    coin.reseed(tree.root()); fri_roots.push_back(tree.root());
    E alpha = coin.draw<E>(); alphas.push_back(alpha);
    u64 gl_inv = finv(root_of_unity(ilog2(Nl))), oinv = finv(offset);
    std::vector<u64> xinv = power_series(gl_inv, rows, oinv);
    std::vector<E> next(rows);
#pragma omp parallel for num_threads(T) schedule(static)
    for (size_t i = 0; i < rows; i++) next[i] = fold_row<E>(&tr[i * F], F, xinv[i], alpha);
    layers.push_back(FriLayerData<E>{std::move(tr), std::move(tree)});
    evals = std::move(next);
  }
  interpolate_poly_with_offset(evals, offset);
  std::vector<E> remainder(evals.begin(), evals.begin() + evals.size() / b);
  Digest rem_commit = hash_elements(remainder);
  coin.reseed(rem_commit);
  st.ms[ST_FRI_LAYERS] = tm.lap();

  // 7 ----- determine_query_positions: grinding (serial smallest nonce) + draw_integers, sort, dedup (A.5)
  u64 nonce = 0;
  for (u64 v = 1;; v++) if (coin.check_leading_zeros(v) >= opt.grinding) { nonce = v; break; }
  std::vector<size_t> positions = coin.draw_integers(opt.num_queries, N, nonce);
  std::sort(positions.begin(), positions.end());
  positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
  st.ms[ST_QUERY_POSITIONS] = tm.lap();

  // 8 ----- build_proof_object (A.10-A.12)
  std::vector<u8> out;
  write_context(out, n, opt, W);
  out.push_back((u8)positions.size());
  { std::vector<u8> cm; cm.insert(cm.end(), trace_tree.root().begin(), trace_tree.root().end());
    cm.insert(cm.end(), ctree.root().begin(), ctree.root().end());
    for (auto& r : fri_roots) cm.insert(cm.end(), r.begin(), r.end());
    cm.insert(cm.end(), rem_commit.begin(), rem_commit.end());
    put_u16(out, cm.size()); put_bytes(out, cm); }
  { std::vector<u8> vals; for (size_t p : positions) for (size_t j = 0; j < W; j++) put_elem(vals, lde[j][p]);
    std::vector<u8> paths = trace_tree.prove_batch(positions).serialize_nodes();
    put_u32(out, vals.size()); put_bytes(out, vals); put_u32(out, paths.size()); put_bytes(out, paths); }
  { std::vector<u8> vals; for (size_t p : positions) for (size_t i = 0; i < K; i++) put_elem(vals, hldes[i][p]);
    std::vector<u8> paths = ctree.prove_batch(positions).serialize_nodes();
    put_u32(out, vals.size()); put_bytes(out, vals); put_u32(out, paths.size()); put_bytes(out, paths); }
  put_u16(out, ood_trace_bytes.size()); put_bytes(out, ood_trace_bytes);
  put_u16(out, ood_eval_bytes.size()); put_bytes(out, ood_eval_bytes);
  out.push_back((u8)layers.size());
  { std::vector<size_t> pos = positions; size_t domain = N;
    for (auto& L : layers) {
      pos = fold_positions(pos, domain, F);
      std::vector<u8> vals; for (size_t p : pos) for (size_t j = 0; j < F; j++) put_elem(vals, L.evals[p * F + j]);
      std::vector<u8> paths = L.tree.prove_batch(pos).serialize_nodes();
      put_u32(out, vals.size()); put_bytes(out, vals); put_u32(out, paths.size()); put_bytes(out, paths);
      domain /= F;
    } }
  { std::vector<u8> rb; put_elems(rb, remainder); put_u16(out, rb.size()); put_bytes(out, rb); }
  out.push_back(0);           // FriProof::num_partitions is serialised as log2(partitions): 0 for the single partition of a serial prover (pinned by running the reference binary, tests/golden/reference_proofs.json)
  put_u64(out, nonce);
  st.ms[ST_BUILD_PROOF] = tm.lap();

  if (times) *times = st;
  if (dbg) {
    dbg->trace_root = trace_tree.root(); dbg->constraint_root = ctree.root(); dbg->fri_roots = fri_roots; dbg->remainder_commitment = rem_commit;
    dbg->tcoef = tcoef; dbg->bcoef = bcoef; dbg->dcoef = dcoef; dbg->alphas = alphas; dbg->ood_frame = frame; dbg->remainder = remainder;
    dbg->z = z; dbg->hz = hz; dbg->nonce = nonce; dbg->positions = positions;
  }
  return out;
}
// the hard-wired normalised BurnMintAir
template <class E>
std::vector<u8> prove(const std::vector<std::vector<F1>>& trace, const PublicInputs& pi, const AirConsts& ac,
                      const ProofOptions& opt, StageTimes* times = nullptr, ProverDebug<E>* dbg = nullptr) {
  if (trace.size() != XFG_TRACE_WIDTH) throw std::runtime_error("trace must have 7 columns");
  return prove<E>(trace, burn_mint_air(pi, ac, trace[0].size()), opt, times, dbg);
}

}  // namespace orc
