// oracle/verifier.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Restates `winterfell::verify::<XfgBurnMintAir, Blake3_256, DefaultRandomCoin>` as called by
// XfgBurnMintVerifier::verify_with_winterfell (src/burn_mint_verifier.rs:265-283) for the normalised
// BurnMintAir, per SURVEY.md A.14.  It shares only field/hash primitives with the oracle prover: every
// protocol value (coefficients, OOD identity, DEEP values, FRI folds) is recomputed from the proof bytes.
#pragma once
#include <algorithm>
#include <string>
#include "air.hpp"
#include "coin.hpp"
#include "merkle.hpp"
#include "ntt.hpp"
#include "prover.hpp"   // fold_positions only

namespace orc {

struct Reader {
  const u8* p; size_t len, pos = 0; bool ok = true;
  Reader(const u8* d, size_t l) : p(d), len(l) {}
  bool need(size_t k) { if (pos + k > len) { ok = false; return false; } return true; }
  u64 uint(int bytes) { if (!need(bytes)) return 0; u64 v = 0; for (int i = bytes - 1; i >= 0; i--) v = (v << 8) | p[pos + i]; pos += bytes; return v; }
  std::vector<u8> bytes(size_t k) { if (!need(k)) return {}; std::vector<u8> r(p + pos, p + pos + k); pos += k; return r; }
  Digest digest() { Digest d{}; if (need(32)) { std::memcpy(d.data(), p + pos, 32); pos += 32; } return d; }
};
template <class E> inline bool read_elems(const std::vector<u8>& b, size_t count, std::vector<E>& out) {
  if (b.size() != count * 8 * E::DEG) return false;
  out.resize(count);
  for (size_t i = 0; i < count; i++) for (int l = 0; l < E::DEG; l++) { u64 v = get_u64(&b[(i * E::DEG + l) * 8]); if (v >= P) return false; out[i].set_limb(l, v); }
  return true;
}
// Lagrange interpolation through (xs[j], ys[j]) evaluated at a (polynom::interpolate_batch + eval on the verifier side)
template <class E> inline E lagrange_eval(const std::vector<E>& xs, const std::vector<E>& ys, E a) {
  E r = E::zero();
  for (size_t j = 0; j < xs.size(); j++) {
    E num = E::one(), den = E::one();
    for (size_t k = 0; k < xs.size(); k++) if (k != j) { num = num * (a - xs[k]); den = den * (xs[j] - xs[k]); }
    r = r + ys[j] * num * den.inv();
  }
  return r;
}

// returns "" when the proof is accepted, otherwise the reason for rejection
template <class E>
std::string verify(const u8* proof, size_t proof_len, AirDef air, const ProofOptions& acceptable, bool burn_mint_pi = false, const PublicInputs* pi = nullptr) {
  const size_t W = air.width;
  Reader rd(proof, proof_len);
  // ---- (1) parse (A.12) ----
  if (rd.uint(1) != W || rd.uint(1) != 0 || rd.uint(1) != 0) return "bad trace layout";
  unsigned lg = (unsigned)rd.uint(1); if (lg < 3 || lg > 32) return "bad trace length";
  if (rd.uint(2) != 0) return "unexpected trace meta";
  if (rd.uint(1) != 8 || rd.uint(8) != P) return "bad field modulus";
  ProofOptions opt; opt.num_queries = (u32)rd.uint(1); opt.blowup = (u32)rd.uint(1); opt.grinding = (u32)rd.uint(1);
  opt.ext = (u32)rd.uint(1); opt.folding = (u32)rd.uint(1); opt.rem_max_deg = (u32)rd.uint(1);
  if (!rd.ok) return "truncated proof";
  if (opt.num_queries != acceptable.num_queries || opt.blowup != acceptable.blowup || opt.grinding != acceptable.grinding ||
      opt.ext != acceptable.ext || opt.folding != acceptable.folding || opt.rem_max_deg != acceptable.rem_max_deg) return "UnacceptableProofOptions";
  if ((int)opt.ext != (E::DEG == 1 ? XFG_EXT_NONE : E::DEG == 2 ? XFG_EXT_QUADRATIC : XFG_EXT_CUBIC)) return "field extension mismatch";
  const size_t n = size_t(1) << lg, b = opt.blowup, N = n * b, F = opt.folding;
  const size_t num_layers = opt.num_fri_layers(N);
  if (burn_mint_pi) air = burn_mint_air(*pi, air.ac, n);      // the burn-mint assertions depend on the trace length
  else { std::string e = validate_air(air, n); if (!e.empty()) return e; }
  const size_t NT = air.num_transition;
  size_t num_unique = rd.uint(1);
  std::vector<u8> cm = rd.bytes(rd.uint(2));
  if (!rd.ok || cm.size() != 32 * (3 + num_layers)) return "bad commitments";
  auto cmd = [&](size_t i) { Digest d; std::memcpy(d.data(), &cm[32 * i], 32); return d; };
  Digest trace_root = cmd(0), constraint_root = cmd(1), rem_commit = cmd(2 + num_layers);
  std::vector<u8> tq_vals = rd.bytes(rd.uint(4)), tq_paths = rd.bytes(rd.uint(4));
  std::vector<u8> cq_vals = rd.bytes(rd.uint(4)), cq_paths = rd.bytes(rd.uint(4));
  std::vector<u8> ood_t = rd.bytes(rd.uint(2)), ood_e = rd.bytes(rd.uint(2));
  if (!rd.ok) return "truncated proof";
  if (rd.uint(1) != num_layers) return "wrong number of FRI layers";
  std::vector<std::vector<u8>> fl_vals(num_layers), fl_paths(num_layers);
  for (size_t l = 0; l < num_layers; l++) { fl_vals[l] = rd.bytes(rd.uint(4)); fl_paths[l] = rd.bytes(rd.uint(4)); }
  std::vector<u8> rem_bytes = rd.bytes(rd.uint(2));
  if (rd.uint(1) != 0) return "bad partition count";   // log2(num_partitions), one partition
  u64 nonce = rd.uint(8);
  if (!rd.ok || rd.pos != proof_len) return "proof length mismatch";

  // ---- (2) replay the transcript ----
  RandomCoin coin(seed_elements(n, opt, W, air.pub_inputs));
  coin.reseed(trace_root);
  std::vector<E> tcoef(NT), bcoef(air.assertions.size());
  for (auto& x : tcoef) x = coin.draw<E>();
  for (auto& x : bcoef) x = coin.draw<E>();
  coin.reseed(constraint_root);
  E z = coin.draw<E>();

  // ---- (3) OOD consistency ----
  if (ood_t.empty() || ood_t[0] != 2) return "bad OOD frame";
  std::vector<E> frame, hz;
  if (!read_elems(std::vector<u8>(ood_t.begin() + 1, ood_t.end()), 2 * W, frame)) return "bad OOD frame";
  const size_t K = air.comp_columns();
  if (!read_elems(ood_e, K, hz)) return "bad OOD evaluations";
  coin.reseed(hash_elements(frame));
  const u64 g_n = root_of_unity(ilog2(n)), g_last = fpow(g_n, n - 1);
  {
    std::vector<E> cur(W), nxt(W), r(NT);
    for (size_t j = 0; j < W; j++) { cur[j] = frame[2 * j]; nxt[j] = frame[2 * j + 1]; }
    eval_air_transition<E>(air, cur.data(), nxt.data(), r.data());
    E t = E::zero(); for (size_t k = 0; k < NT; k++) t = t + tcoef[k] * r[k];
    E zn = epow(z, (u64)n);
    E result = t * (z - E::from_base(g_last)) * (zn - E::one()).inv();
    const std::vector<Assertion>& asr = air.assertions;
    for (size_t k = 0; k < asr.size();) {       // one boundary group per distinct step: sum of its terms over (z - g^step)
      E bsum = E::zero(); const size_t step = asr[k].step;
      for (; k < asr.size() && asr[k].step == step; k++) bsum = bsum + bcoef[k] * (cur[asr[k].column] - E::from_base(asr[k].value));
      result = result + bsum * (z - E::from_base(fpow(g_n, step))).inv();
    }
    // the composition polynomial at z from its K columns: sum_i z^(i n) H_i(z)
    E hsum = E::zero(), zp = E::one();
    for (size_t i = 0; i < K; i++) { hsum = hsum + hz[i] * zp; zp = zp * zn; }
    if (result != hsum) return "InconsistentOodConstraintEvaluations";
  }
  coin.reseed(hash_elements(hz));

  // ---- (4) DEEP coefficients, FRI alphas ----
  std::vector<E> dcoef(W + K); for (auto& x : dcoef) x = coin.draw<E>();
  std::vector<E> alphas;
  for (size_t l = 0; l < num_layers; l++) { coin.reseed(cmd(2 + l)); alphas.push_back(coin.draw<E>()); }
  coin.reseed(rem_commit);

  // ---- (5) proof of work + query positions ----
  if (coin.check_leading_zeros(nonce) < opt.grinding) return "QuerySeedProofOfWorkVerificationFailed";
  std::vector<size_t> positions = coin.draw_integers(opt.num_queries, N, nonce);
  std::sort(positions.begin(), positions.end());
  positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
  if (positions.size() != num_unique) return "NumberOfQueriesMismatch";

  // ---- (6) trace / constraint openings ----
  std::vector<F1> trows; std::vector<E> crows;
  if (!read_elems(tq_vals, positions.size() * W, trows)) return "bad trace query values";
  if (!read_elems(cq_vals, positions.size() * K, crows)) return "bad constraint query values";
  {
    std::vector<Digest> lv(positions.size());
    for (size_t i = 0; i < positions.size(); i++) lv[i] = hash_elements(&trows[i * W], W);
    BatchMerkleProof bp; Digest root;
    if (!BatchMerkleProof::deserialize(tq_paths.data(), tq_paths.size(), lv, ilog2(N), bp)) return "bad trace query paths";
    if (!bp.get_root(positions, root) || root != trace_root) return "TraceQueryDoesNotMatchCommitment";
    for (size_t i = 0; i < positions.size(); i++) lv[i] = hash_elements(&crows[i * K], K);
    BatchMerkleProof cp;
    if (!BatchMerkleProof::deserialize(cq_paths.data(), cq_paths.size(), lv, ilog2(N), cp)) return "bad constraint query paths";
    if (!cp.get_root(positions, root) || root != constraint_root) return "ConstraintQueryDoesNotMatchCommitment";
  }

  // ---- (7) DEEP composition at the queried points ----
  const u64 g_N = root_of_unity(ilog2(N));
  E zg = z.mul_base(g_n);
  std::vector<E> evaluations(positions.size());
  for (size_t i = 0; i < positions.size(); i++) {
    E x = E::from_base(fmul(XFG_GENERATOR, fpow(g_N, positions[i])));
    E i1 = (x - z).inv(), i2 = (x - zg).inv(), acc = E::zero();
    for (size_t j = 0; j < W; j++) {
      E t = E::from_base(trows[i * W + j].v);
      acc = acc + dcoef[j] * ((t - frame[2 * j]) * i1 + (t - frame[2 * j + 1]) * i2);
    }
    for (size_t k = 0; k < K; k++) acc = acc + dcoef[W + k] * (crows[i * K + k] - hz[k]) * i1;
    evaluations[i] = acc;
  }

  // ---- (8) FRI ----
  {
    std::vector<size_t> pos = positions; size_t domain = N; u64 gen = g_N;
    size_t max_deg_plus_1 = n;   // DEEP polynomial has degree < n
    std::vector<u64> folding_roots(F); { u64 wF = root_of_unity(ilog2(F)); for (size_t j = 0; j < F; j++) folding_roots[j] = fpow(wF, j); }
    for (size_t l = 0; l < num_layers; l++) {
      std::vector<size_t> folded = fold_positions(pos, domain, F);
      std::vector<E> vals; if (!read_elems(fl_vals[l], folded.size() * F, vals)) return "bad FRI layer values";
      std::vector<Digest> lv(folded.size()); for (size_t i = 0; i < folded.size(); i++) lv[i] = hash_elements(&vals[i * F], F);
      BatchMerkleProof bp; Digest root;
      if (!BatchMerkleProof::deserialize(fl_paths[l].data(), fl_paths[l].size(), lv, ilog2(domain / F), bp)) return "bad FRI layer paths";
      if (!bp.get_root(folded, root) || root != cmd(2 + l)) return "LayerCommitmentMismatch";
      size_t row_len = domain / F;
      for (size_t i = 0; i < pos.size(); i++) {
        size_t idx = std::find(folded.begin(), folded.end(), pos[i] % row_len) - folded.begin();
        if (vals[idx * F + pos[i] / row_len] != evaluations[i]) return "InvalidLayerFolding";
      }
      std::vector<E> nxt(folded.size());
      for (size_t i = 0; i < folded.size(); i++) {
        u64 xe = fmul(fpow(gen, folded[i]), XFG_GENERATOR);    // constant offset at every layer (A.10, D)
        std::vector<E> xs(F), ys(vals.begin() + i * F, vals.begin() + (i + 1) * F);
        for (size_t j = 0; j < F; j++) xs[j] = E::from_base(fmul(xe, folding_roots[j]));
        nxt[i] = lagrange_eval(xs, ys, alphas[l]);
      }
      if (max_deg_plus_1 % F != 0) return "DegreeTruncation";
      evaluations = nxt; pos = folded; gen = fpow(gen, F); max_deg_plus_1 /= F; domain /= F;
    }
    std::vector<E> rem; if (rem_bytes.size() % (8 * E::DEG) || !read_elems(rem_bytes, rem_bytes.size() / (8 * E::DEG), rem)) return "bad remainder";
    if (hash_elements(rem) != rem_commit) return "RemainderCommitmentMismatch";
    if (rem.size() > max_deg_plus_1) return "RemainderDegreeMismatch";
    for (size_t i = 0; i < pos.size(); i++) {
      E x = E::from_base(fmul(XFG_GENERATOR, fpow(gen, pos[i])));
      if (eval_poly<E, E>(rem, x) != evaluations[i]) return "InvalidRemainderFolding";
    }
  }
  return "";
}
// the hard-wired normalised BurnMintAir
template <class E>
std::string verify(const u8* proof, size_t proof_len, const PublicInputs& pi, const AirConsts& ac, const ProofOptions& acceptable) {
  AirDef a; a.burn_mint = true; a.ac = ac; a.width = XFG_TRACE_WIDTH;
  return verify<E>(proof, proof_len, a, acceptable, true, &pi);
}

}  // namespace orc
