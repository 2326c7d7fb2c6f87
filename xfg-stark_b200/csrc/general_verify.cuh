// general_verify.cuh — batch verification for the general option space: any `ProofOptions` (blowup 2..128, folding 2/4/8/16, remainder degree 0..255,
// None / Quadratic / Cubic) and any compiled AIR (transition degrees up to 9, i.e. up to 8 composition columns).
//
// Replaces `winterfell::verify::<_, Blake3_256, DefaultRandomCoin>` (src/burn_mint_verifier.rs:265-283) and the sequential loop around it
// (`BatchBurnMintVerifier::verify_batch`, :371-408) where the tuned verifier (verify.cu: one thread block per proof, options 8 / 8, burn-mint AIR) does
// not apply.  ONE THREAD PER PROOF: the body below walks the proof bytes serially in the order winter-verifier runs its checks, so
// the first failing check is the one the serial verifier names; a batch is verified by as many threads as it has proofs (~4000 dependent BLAKE3
// compressions per proof: a few ms of latency, amortised over the batch).  Like the prover bodies (general_bodies.cuh) it is host+device code that
// tests/host_emul also runs on a CPU-only box against the oracle verifier's verdicts.
#pragma once
#include "../../include/xfg_stark.h"
#include "general_bodies.cuh"

namespace xfg {

struct GoVerifyRec {                 // one proof of the batch (host-prepared)
  u64 proof_off; u32 proof_len; u32 prog_off;      // byte offsets into the staged proofs / compiled programs
  u32 seed_count, host_status;                     // host_status != 0: rejected on the host (too short to read the trace length, bad AIR): no device work
  u64 seed_limbs[MAX_SEED_LIMBS];                  // Context::to_elements || public inputs (A.4)
};
struct GoVerifyWork {                // per-proof scratch in global memory
  u64 coef[GEN_MAX_CONSTRAINTS + GEN_MAX_ASSERTIONS][GO_MAX_EXT], dcoef[GEN_MAX_WIDTH + GO_MAX_CE][GO_MAX_EXT], alphas[GO_MAX_LAYERS][GO_MAX_EXT];
  u64 evals[256][GO_MAX_EXT], nxt[256][GO_MAX_EXT];
  u32 positions[256], pos[256], folded[256], norm[256], node_off[256];
  u64 cur_idx[256], nxt_idx[256];
  Digest lv[256], cur_dig[256], nxt_dig[256];
  u8 node_cnt[256], ptr[256];
};

XFG_HD u64 go_rd(const u8* p, int bytes) { u64 v = 0; for (int i = bytes - 1; i >= 0; i--) v = (v << 8) | p[i]; return v; }
XFG_HD Digest go_rd_digest(const u8* p) { Digest d; for (int i = 0; i < 8; i++) d.w[i] = (u32)go_rd(p + 4 * i, 4); return d; }
XFG_HD bool go_digest_eq(const Digest& a, const Digest& b) { bool e = true; for (int i = 0; i < 8; i++) e &= a.w[i] == b.w[i]; return e; }
// count elements of D limbs at p, every limb canonical
template <int D> XFG_HD bool go_canonical(const u8* p, size_t count) { for (size_t i = 0; i < count * D; i++) if (go_rd(p + 8 * i, 8) >= GL_P) return false; return true; }
template <int D> XFG_HD Ext<D> go_rd_ext(const u8* p) { Ext<D> r; for (int l = 0; l < D; l++) r.set_limb(l, go_rd(p + 8 * l, 8)); return r; }
template <int D> XFG_HD bool go_ext_eq(const Ext<D>& a, const Ext<D>& b) { bool e = true; for (int l = 0; l < D; l++) e &= a.limb(l) == b.limb(l); return e; }

// winter-crypto BatchMerkleProof::{deserialize, get_root} (A.11) for `cnt` leaves lv[] at indexes idx[] (distinct, < 2^depth) against serialised
// nodes at p[0..len).  Returns 0 = root computed, 1 = malformed serialisation (the oracle's "bad ... paths"), 2 = get_root fails
XFG_HD int go_batch_root(GoVerifyWork& w, const u32* idx, u32 cnt, const u8* p, size_t len, u32 depth, Digest& root) {
  // deserialize: u8 count, then per vector u8 len + digests, exact length
  if (len < 1) return 1;
  size_t pos = 1; const u32 nv = p[0];
  for (u32 i = 0; i < nv; i++) {
    if (pos >= len) return 1;
    const u32 k = p[pos++]; if (pos + 32 * (size_t)k > len) return 1;
    w.node_off[i] = (u32)pos; w.node_cnt[i] = (u8)k; pos += 32 * (size_t)k;
  }
  if (pos != len) return 1;
  // get_root
  if (cnt == 0 || cnt > 255) return 2;
  for (u32 i = 0; i < cnt; i++) { if ((u64)idx[i] >> depth) return 2; for (u32 j = 0; j < i; j++) if (idx[j] == idx[i]) return 2; }      // map_indexes: range, duplicates
  u32 nn = 0;                                                   // normalize_indexes: sorted unique (index & ~1)
  for (u32 i = 0; i < cnt; i++) {
    const u32 v = idx[i] & ~1u; u32 at = nn; bool dup = false;
    for (u32 j = 0; j < nn; j++) if (w.norm[j] == v) dup = true;
    if (dup) continue;
    while (at > 0 && w.norm[at - 1] > v) { w.norm[at] = w.norm[at - 1]; at--; }
    w.norm[at] = v; nn++;
  }
  if (nn != nv) return 2;
  auto find = [&](u32 index) -> int { for (u32 i = 0; i < cnt; i++) if (idx[i] == index) return (int)i; return -1; };
  auto node = [&](u32 i, u32 k) { return go_rd_digest(p + w.node_off[i] + 32 * (size_t)k); };
  const u64 offset = u64(1) << depth;
  u32 nc = 0;
  for (u32 i = 0; i < nn; i++) {
    const u32 index = w.norm[i]; Digest b0, b1; const int i1 = find(index), i2 = find(index + 1);
    if (i1 >= 0) {
      b0 = w.lv[i1];
      if (i2 >= 0) { b1 = w.lv[i2]; w.ptr[i] = 0; }
      else { if (w.node_cnt[i] == 0) return 2; b1 = node(i, 0); w.ptr[i] = 1; }
    } else {
      if (w.node_cnt[i] == 0) return 2;
      b0 = node(i, 0);
      if (i2 < 0) return 2;
      b1 = w.lv[i2]; w.ptr[i] = 1;
    }
    w.cur_idx[nc] = (offset + index) >> 1; w.cur_dig[nc] = go_merge(b0, b1); nc++;
  }
  for (u32 d = 1; d < depth; d++) {
    u32 nnext = 0, i = 0;
    while (i < nc) {
      const u64 node_index = w.cur_idx[i], sib_index = node_index ^ 1; Digest sib; const u32 self = i;
      if (i + 1 < nc && w.cur_idx[i + 1] == sib_index) { sib = w.cur_dig[i + 1]; i += 1; }
      else { const u32 q = w.ptr[self]; if (w.node_cnt[self] <= q) return 2; sib = node(self, q); w.ptr[self] += 1; }      // nodes[i] / ptr[i] are indexed by the position in the current list (A.11)
      const Digest parent = (node_index & 1) ? go_merge(sib, w.cur_dig[self]) : go_merge(w.cur_dig[self], sib);
      w.nxt_idx[nnext] = node_index >> 1; w.nxt_dig[nnext] = parent; nnext++;
      i += 1;
    }
    for (u32 j = 0; j < nnext; j++) { w.cur_idx[j] = w.nxt_idx[j]; w.cur_dig[j] = w.nxt_dig[j]; }
    nc = nnext;
  }
  if (nc != 1 || w.cur_idx[0] != 1) return 2;
  root = w.cur_dig[0]; return 0;
}

template <int D> struct GoVerify {
  const GoVerifyRec* recs; const u8* bytes; const u8* progs; GoVerifyWork* work; xfg_options opt; int* results;
  XFG_HD void operator()(size_t t) const {
    const GoVerifyRec& r = recs[t];
    results[t] = r.host_status ? (int)r.host_status : run(r, bytes + r.proof_off, r.proof_len, *reinterpret_cast<const GenProgram*>(progs + r.prog_off), work[t]);
  }
  XFG_HD int run(const GoVerifyRec& rec, const u8* p, size_t len, const GenProgram& prog, GoVerifyWork& w) const {
    const u32 W = prog.width, NT = prog.num_constraints, A = prog.num_assertions, K = prog.max_degree > 2 ? prog.max_degree - 1 : 1;
    size_t at = 0; bool ok = true;
    auto rd = [&](int nb) -> u64 { if (at + nb > len) { ok = false; return 0; } const u64 v = go_rd(p + at, nb); at += nb; return v; };
    auto skip = [&](size_t k) -> size_t { if (at + k > len) { ok = false; return 0; } const size_t s = at; at += k; return s; };
    // ---- (1) parse (A.12) ----
    if (rd(1) != W || rd(1) != 0 || rd(1) != 0) return XFG_VERIFY_MALFORMED;
    const u32 lg = (u32)rd(1); if (lg < 3 || lg > 27) return XFG_VERIFY_MALFORMED;
    if (rd(2) != 0) return XFG_VERIFY_MALFORMED;
    if (rd(1) != 8 || rd(8) != GL_P) return XFG_VERIFY_MALFORMED;
    u32 po[6]; for (int i = 0; i < 6; i++) po[i] = (u32)rd(1);
    if (!ok) return XFG_VERIFY_MALFORMED;
    if (po[0] != opt.num_queries || po[1] != opt.blowup_factor || po[2] != opt.grinding_factor || po[3] != opt.field_extension || po[4] != opt.fri_folding_factor ||
        po[5] != opt.fri_remainder_max_degree) return XFG_VERIFY_UNACCEPTABLE_OPTIONS;
    u32 lb = 0; while ((1u << lb) < opt.blowup_factor) lb++;
    u32 lf = 0; while ((1u << lf) < opt.fri_folding_factor) lf++;
    const u32 lN = lg + lb, F = 1u << lf; if (lN > 31) return XFG_VERIFY_MALFORMED;
    const u64 n = u64(1) << lg, N = u64(1) << lN;
    u32 L = 0; { u32 l = lN; const u64 mx = (u64)(opt.fri_remainder_max_degree + 1) * opt.blowup_factor;
      while ((u64(1) << l) > mx) { if (l < lf + 1 || L == (u32)GO_MAX_LAYERS) return XFG_VERIFY_MALFORMED; l -= lf; L++; } }
    const u32 num_unique = (u32)rd(1);
    const size_t cl = (size_t)rd(2), cat = skip(cl);
    if (!ok || cl != 32 * (size_t)(3 + L)) return XFG_VERIFY_MALFORMED;
    auto cmd = [&](u32 i) { return go_rd_digest(p + cat + 32 * (size_t)i); };
    const Digest trace_root = cmd(0), constraint_root = cmd(1), rem_commit = cmd(2 + L);
    size_t sec_at[4], sec_len[4];
    for (int i = 0; i < 4; i++) { sec_len[i] = (size_t)rd(4); sec_at[i] = skip(sec_len[i]); }      // trace values, trace paths, constraint values, constraint paths
    const size_t otl = (size_t)rd(2), otat = skip(otl), oel = (size_t)rd(2), oeat = skip(oel);
    if (!ok) return XFG_VERIFY_MALFORMED;
    if (rd(1) != L) return XFG_VERIFY_MALFORMED;
    size_t fv_at[GO_MAX_LAYERS], fv_len[GO_MAX_LAYERS], fp_at[GO_MAX_LAYERS], fp_len[GO_MAX_LAYERS];
    for (u32 l = 0; l < L; l++) { fv_len[l] = (size_t)rd(4); fv_at[l] = skip(fv_len[l]); fp_len[l] = (size_t)rd(4); fp_at[l] = skip(fp_len[l]); }
    const size_t rl = (size_t)rd(2), rat = skip(rl);
    if (rd(1) != 0) return XFG_VERIFY_MALFORMED;               // FriProof::num_partitions as log2: one partition
    const u64 nonce = rd(8);
    if (!ok || at != len) return XFG_VERIFY_MALFORMED;

    // ---- (2) replay the transcript ----
    const u64* sl = rec.seed_limbs;
    GoCoin c; c.seed = go_hash_stream((int)rec.seed_count, [sl](int i) { return sl[i]; }); c.counter = 0;
    go_reseed(c, trace_root);
    for (u32 i = 0; i < NT + A; i++) { u64 v[GO_MAX_EXT] = {0, 0, 0}; if (!go_draw<D>(c, v)) return XFG_VERIFY_COIN; for (int l = 0; l < GO_MAX_EXT; l++) w.coef[i][l] = v[l]; }
    go_reseed(c, constraint_root);
    Ext<D> z; { u64 v[GO_MAX_EXT] = {0, 0, 0}; if (!go_draw<D>(c, v)) return XFG_VERIFY_COIN; z = go_ld<D>(v); }

    // ---- (3) OOD consistency ----
    if (otl == 0 || p[otat] != 2) return XFG_VERIFY_MALFORMED;
    if (otl != 1 + (size_t)2 * W * D * 8 || !go_canonical<D>(p + otat + 1, (size_t)2 * W)) return XFG_VERIFY_MALFORMED;
    if (oel != (size_t)K * D * 8 || !go_canonical<D>(p + oeat, K)) return XFG_VERIFY_MALFORMED;
    const u8* fr = p + otat + 1; const u8* hzp = p + oeat;
    go_reseed(c, go_hash_stream((int)(2 * W * D), [fr](int li) { return go_rd(fr + 8 * (size_t)li, 8); }));
    const u64 g_n = gl_root_of_unity(lg), g_last = gl_pow(g_n, n - 1);
    {
      // Air::evaluate_transition on the OOD frame, in the extension field: the compiled program over Ext<D> slots
      Ext<D> slot[GEN_MAX_SLOTS]; Ext<D> tsum;
      for (u32 i = 0; i < prog.num_instr; i++) {
        const GenInstr in = prog.code[i]; const u32 op = in.w0 & 15u, dst = in.w0 >> 8;
        Ext<D> v[2];
        for (int o = 0; o < 2; o++) {
          const u32 kind = (in.w0 >> (4 + 2 * o)) & 3u, idx = o ? in.w1 >> 16 : in.w1 & 0xFFFFu;
          v[o] = kind == GK_SLOT ? slot[idx] : kind == GK_CONST ? Ext<D>(prog.constants[idx]) : go_rd_ext<D>(fr + 8 * (size_t)D * (2 * idx + (kind == GK_CUR ? 0 : 1)));
          if (op == GOP_OUT) break;
        }
        if (op == GOP_OUT) { tsum = tsum + go_ld<D>(w.coef[dst]) * v[0]; continue; }
        slot[dst] = op == GOP_MUL ? v[0] * v[1] : op == GOP_ADD ? v[0] + v[1] : v[0] - v[1];
      }
      const Ext<D> zn = go_pow<D>(z, n);
      Ext<D> result = tsum * (z - Ext<D>(g_last)) * ext_inv(zn - Ext<D>(1));
      for (u32 k = 0; k < A;) {         // one boundary group per distinct step: sum of its terms over (z - g^step)
        Ext<D> bsum; const u32 g = prog.asr[k].group;
        for (; k < A && prog.asr[k].group == g; k++) bsum = bsum + go_ld<D>(w.coef[NT + k]) * (go_rd_ext<D>(fr + 8 * (size_t)D * (2 * prog.asr[k].column)) - Ext<D>(prog.asr[k].value));
        result = result + bsum * ext_inv(z - Ext<D>(prog.group_point[g]));
      }
      Ext<D> hsum, zp(1);              // the composition polynomial at z from its K columns: sum_i z^(i n) H_i(z)
      for (u32 i = 0; i < K; i++) { hsum = hsum + go_rd_ext<D>(hzp + 8 * (size_t)D * i) * zp; zp = zp * zn; }
      if (!go_ext_eq<D>(result, hsum)) return XFG_VERIFY_INCONSISTENT_OOD;
    }
    go_reseed(c, go_hash_stream((int)(K * D), [hzp](int li) { return go_rd(hzp + 8 * (size_t)li, 8); }));

    // ---- (4) DEEP coefficients, FRI alphas ----
    for (u32 i = 0; i < W + K; i++) { u64 v[GO_MAX_EXT] = {0, 0, 0}; if (!go_draw<D>(c, v)) return XFG_VERIFY_COIN; for (int l = 0; l < GO_MAX_EXT; l++) w.dcoef[i][l] = v[l]; }
    for (u32 l = 0; l < L; l++) { go_reseed(c, cmd(2 + l)); u64 v[GO_MAX_EXT] = {0, 0, 0}; if (!go_draw<D>(c, v)) return XFG_VERIFY_COIN; for (int q = 0; q < GO_MAX_EXT; q++) w.alphas[l][q] = v[q]; }
    go_reseed(c, rem_commit);

    // ---- (5) proof of work + query positions ----
    { const Digest d = go_merge_int(c.seed, nonce); const u64 head = (u64)d.w[0] | ((u64)d.w[1] << 32);
      const u32 tz = head == 0 ? 64 : [&] { u32 k = 0; while (!((head >> k) & 1)) k++; return k; }();
      if (tz < opt.grinding_factor) return XFG_VERIFY_POW_FAILED; }
    u32 cnt = 0;
    { c.seed = go_merge_int(c.seed, nonce); c.counter = 0;
      for (u32 i = 0; i < opt.num_queries; i++) {
        c.counter += 1; const Digest d = go_merge_int(c.seed, c.counter);
        const u32 v = (u32)(((u64)d.w[0] | ((u64)d.w[1] << 32)) & (N - 1));
        u32 q = cnt; bool dup = false; for (u32 j = 0; j < cnt; j++) if (w.positions[j] == v) dup = true;
        if (dup) continue;
        while (q > 0 && w.positions[q - 1] > v) { w.positions[q] = w.positions[q - 1]; q--; }
        w.positions[q] = v; cnt++;
      } }
    if (cnt != num_unique) return XFG_VERIFY_NUM_QUERIES_MISMATCH;

    // ---- (6) trace / constraint openings ----
    if (sec_len[0] != (size_t)cnt * W * 8 || !go_canonical<1>(p + sec_at[0], (size_t)cnt * W)) return XFG_VERIFY_MALFORMED;
    if (sec_len[2] != (size_t)cnt * K * D * 8 || !go_canonical<D>(p + sec_at[2], (size_t)cnt * K)) return XFG_VERIFY_MALFORMED;
    const u8* tv = p + sec_at[0]; const u8* cv = p + sec_at[2];
    {
      Digest root;
      for (u32 i = 0; i < cnt; i++) { const u8* row = tv + (size_t)i * W * 8; w.lv[i] = go_hash_stream((int)W, [row](int li) { return go_rd(row + 8 * (size_t)li, 8); }); }
      int rc = go_batch_root(w, w.positions, cnt, p + sec_at[1], sec_len[1], lN, root);
      if (rc == 1) return XFG_VERIFY_MALFORMED;
      if (rc || !go_digest_eq(root, trace_root)) return XFG_VERIFY_TRACE_QUERY_MISMATCH;
      for (u32 i = 0; i < cnt; i++) { const u8* row = cv + (size_t)i * K * D * 8; w.lv[i] = go_hash_stream((int)(K * D), [row](int li) { return go_rd(row + 8 * (size_t)li, 8); }); }
      rc = go_batch_root(w, w.positions, cnt, p + sec_at[3], sec_len[3], lN, root);
      if (rc == 1) return XFG_VERIFY_MALFORMED;
      if (rc || !go_digest_eq(root, constraint_root)) return XFG_VERIFY_CONSTRAINT_QUERY_MISMATCH;
    }

    // ---- (7) DEEP composition at the queried points ----
    const u64 g_N = gl_root_of_unity(lN);
    const Ext<D> zg = mul_base(z, g_n);
    for (u32 i = 0; i < cnt; i++) {
      const Ext<D> x(gl_mul(XFG_GENERATOR, gl_pow(g_N, w.positions[i])));
      const Ext<D> i1 = ext_inv(x - z), i2 = ext_inv(x - zg); Ext<D> acc;
      for (u32 j = 0; j < W; j++) {
        const Ext<D> tj(go_rd(tv + ((size_t)i * W + j) * 8, 8));
        acc = acc + go_ld<D>(w.dcoef[j]) * ((tj - go_rd_ext<D>(fr + 8 * (size_t)D * (2 * j))) * i1 + (tj - go_rd_ext<D>(fr + 8 * (size_t)D * (2 * j + 1))) * i2);
      }
      for (u32 k = 0; k < K; k++) acc = acc + go_ld<D>(w.dcoef[W + k]) * (go_rd_ext<D>(cv + ((size_t)i * K + k) * D * 8) - go_rd_ext<D>(hzp + 8 * (size_t)D * k)) * i1;
      go_st<D>(w.evals[i], acc);
    }

    // ---- (8) FRI ----
    u32 np = cnt; for (u32 i = 0; i < cnt; i++) w.pos[i] = w.positions[i];
    u32 ldom = lN; u64 gen = g_N; u64 max_deg_plus_1 = n;        // the DEEP polynomial has degree < n
    const u64 wfi = gl_inv(gl_root_of_unity(lf)), f_inv = gl_inv(F), inv7 = gl_inv(XFG_GENERATOR);
    for (u32 l = 0; l < L; l++) {
      const u32 lrow = ldom - lf; const u32 row_mask = (u32)((u64(1) << lrow) - 1);
      u32 nf = 0;                                                  // fold_positions: pos mod row_len, first occurrence kept
      for (u32 i = 0; i < np; i++) { const u32 v = w.pos[i] & row_mask; bool dup = false; for (u32 j = 0; j < nf; j++) if (w.folded[j] == v) dup = true; if (!dup) w.folded[nf++] = v; }
      if (fv_len[l] != (size_t)nf * F * D * 8 || !go_canonical<D>(p + fv_at[l], (size_t)nf * F)) return XFG_VERIFY_MALFORMED;
      const u8* vals = p + fv_at[l];
      for (u32 i = 0; i < nf; i++) { const u8* row = vals + (size_t)i * F * D * 8; w.lv[i] = go_hash_stream((int)(F * D), [row](int li) { return go_rd(row + 8 * (size_t)li, 8); }); }
      Digest root; const int rc = go_batch_root(w, w.folded, nf, p + fp_at[l], fp_len[l], lrow, root);
      if (rc == 1) return XFG_VERIFY_MALFORMED;
      if (rc || !go_digest_eq(root, cmd(2 + l))) return XFG_VERIFY_FRI_LAYER_COMMITMENT_MISMATCH;
      for (u32 i = 0; i < np; i++) {
        u32 idx = 0; while (idx < nf && w.folded[idx] != (w.pos[i] & row_mask)) idx++;
        if (!go_ext_eq<D>(go_rd_ext<D>(vals + ((size_t)idx * F + (w.pos[i] >> lrow)) * D * 8), go_ld<D>(w.evals[i]))) return XFG_VERIFY_FRI_INVALID_LAYER_FOLDING;
      }
      const Ext<D> alpha = go_ld<D>(w.alphas[l]);
      for (u32 i = 0; i < nf; i++) {      // the degree < F interpolant of the row over x w_F^j at alpha: inverse DFT + Horner in alpha / x (as GoFriFold)
        const u64 xinv = gl_mul(inv7, gl_inv(gl_pow(gen, w.folded[i])));
        const Ext<D> beta = mul_base(alpha, xinv); Ext<D> acc;
        for (u32 kk = F; kk-- > 0;) {
          Ext<D> ck;
          for (u32 j = 0; j < F; j++) ck = ck + mul_base(go_rd_ext<D>(vals + ((size_t)i * F + j) * D * 8), gl_pow(wfi, (u64)((j * kk) & (F - 1))));
          acc = acc * beta + ck;
        }
        go_st<D>(w.nxt[i], mul_base(acc, f_inv));
      }
      if (max_deg_plus_1 % F != 0) return XFG_VERIFY_FRI_DEGREE_TRUNCATION;
      for (u32 i = 0; i < nf; i++) { for (int q = 0; q < GO_MAX_EXT; q++) w.evals[i][q] = w.nxt[i][q]; w.pos[i] = w.folded[i]; }
      np = nf; gen = gl_pow(gen, F); max_deg_plus_1 /= F; ldom = lrow;
    }
    if (rl % (8 * (size_t)D) || !go_canonical<D>(p + rat, rl / (8 * (size_t)D))) return XFG_VERIFY_MALFORMED;
    const size_t rn = rl / (8 * (size_t)D); const u8* rem = p + rat;
    if (!go_digest_eq(go_hash_stream((int)(rn * D), [rem](int li) { return go_rd(rem + 8 * (size_t)li, 8); }), rem_commit)) return XFG_VERIFY_FRI_REMAINDER_COMMITMENT_MISMATCH;
    if (rn > max_deg_plus_1) return XFG_VERIFY_FRI_REMAINDER_DEGREE_MISMATCH;
    for (u32 i = 0; i < np; i++) {
      const Ext<D> x(gl_mul(XFG_GENERATOR, gl_pow(gen, w.pos[i]))); Ext<D> acc;
      for (size_t k = rn; k-- > 0;) acc = acc * x + go_rd_ext<D>(rem + k * D * 8);
      if (!go_ext_eq<D>(acc, go_ld<D>(w.evals[i]))) return XFG_VERIFY_FRI_INVALID_REMAINDER_FOLDING;
    }
    return XFG_VERIFY_OK;
  }
};

}  // namespace xfg
