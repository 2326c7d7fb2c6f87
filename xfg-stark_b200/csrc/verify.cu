// verify.cu — batch verification of burn-mint STARK proofs for sm_100a: one thread block per proof.
//
// Replaces the sequential loop of `BatchBurnMintVerifier` (src/burn_mint_verifier.rs:371-408) around
// `XfgBurnMintVerifier::verify_with_winterfell` -> `winterfell::verify` (src/burn_mint_verifier.rs:265-283) for the normalised
// BurnMintAir.  The host only walks the proof's length prefixes (verify_api.inc) and uploads the raw bytes;
// every hash, transcript draw, canonicity check and field operation of the verifier runs here, reading the unaligned bytes directly:
//   warp 0      replays the Fiat-Shamir transcript (coin.cuh: lanes hash candidate counters in parallel),
//   warp 1      checks the out-of-domain constraint identity,
//   all threads hash the opened rows, rebuild the BatchMerkleProof roots level by level (one block scan per level), recompute the
//               DEEP composition at the queried points and check every FRI fold and the remainder.
// The checks run in the order of winterfell::verify, so the first failing check names the same error.
#include "verify.cuh"
#include "coin.cuh"
#include "launch.cuh"
#include "../../include/xfg_stark.h"

namespace xfg {

static constexpr int VT = 256;   // threads per proof; >= XFG_MAX_QUERIES

struct VShared {
  Digest digA[VT], digB[VT], leaf[VT];
  u32 idxA[VT], idxB[VT];
  u32 pos[VT], fpos[VT], tmp[VT], slot[VT];
  u64 evals[VT][2], nxt[VT][2];
  u64 coef[XFG_NUM_TRANSITION + XFG_NUM_ASSERTIONS][2], z[1][2], dcoef[XFG_TRACE_WIDTH + 1][2], alphas[MAX_LAYERS][2];
  u32 wsum[VT / 32];
  u32 npos, nf;
  int status;
};

template <int D> __device__ __forceinline__ Ext<D> lde2(const u64* p) { return Ext<D>(p[0], p[1]); }
template <int D> __device__ __forceinline__ Ext<D> sub_base(Ext<D> a, u64 b) { return a - Ext<D>::from_base(b); }
template <int D> __device__ __forceinline__ bool ext_eq(Ext<D> a, Ext<D> b) { return is_zero(a - b); }
__device__ __forceinline__ bool dig_eq(const Digest& a, const Digest& b) { u32 x = 0; for (int i = 0; i < 8; i++) x |= a.w[i] ^ b.w[i]; return x == 0; }
// little-endian u64 at an arbitrary byte offset (two aligned loads + funnel shift; the batch buffer is padded by 8 bytes)
__device__ __forceinline__ u64 ld_u64(const u8* __restrict__ base, size_t off) {
  const size_t a = (size_t)(base + off); const u64* p = reinterpret_cast<const u64*>(a & ~size_t(7)); const u32 sh = (u32)(a & 7) * 8;
  const u64 lo = p[0]; if (sh == 0) return lo;
  return (lo >> sh) | (p[1] << (64 - sh));
}
__device__ __forceinline__ Digest ld_digest(const u8* __restrict__ base, size_t off) {
  Digest d;
#pragma unroll
  for (int i = 0; i < 4; i++) { const u64 v = ld_u64(base, off + 8 * i); d.w[2 * i] = (u32)v; d.w[2 * i + 1] = (u32)(v >> 32); }
  return d;
}
// The XFG_VERIFY_* codes are numbered in the order winterfell::verify runs its checks, so when concurrent checks fail (warp 0 and
// warp 1, or several threads of one phase) the smallest code is the one the serial verifier would have reported.
__device__ __forceinline__ void vfail(VShared& sh, int code) { const int old = atomicCAS(&sh.status, 0, code); if (old != 0 && code < old) atomicMin(&sh.status, code); }

// exclusive prefix sum of `v` over the block (all VT threads call); total = block sum
__device__ __forceinline__ u32 block_excl_sum(VShared& sh, u32 v, u32& total) {
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u32 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= (u32)o) inc += t; }
  if (lane == 31) sh.wsum[w] = inc;
  __syncthreads();
  u32 base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < VT / 32; i++) { const u32 s = sh.wsum[i]; if ((u32)i < w) base += s; tot += s; }
  __syncthreads();
  total = tot;
  return base + inc - v;
}

// BatchMerkleProof::get_root (A.11): `cnt` distinct leaf indexes plist[] in any order, their leaf digests in sh.leaf[] (same
// order), node vectors of the opening in `data`.  All threads call; returns true (block-uniform) with the root in `root`.
__device__ bool batch_root(VShared& sh, const u32* plist, u32 cnt, u32 depth, const VerifyOpening& op, const u8* __restrict__ pb, const uint2* __restrict__ vec_index, Digest& root) {
  const u32 t = threadIdx.x;
  __shared__ int bad;
  if (t == 0) bad = 0;
  __syncthreads();
  if (cnt == 0 || cnt > XFG_MAX_QUERIES) return false;
  // map_indexes + normalize_indexes: sort (rank), reject duplicates / out-of-range, group into sibling pairs
  if (t < cnt) {
    const u32 key = plist[t]; u32 rank = 0; bool dup = false;
    for (u32 j = 0; j < cnt; j++) { const u32 o = plist[j]; rank += o < key; dup |= (o == key && j != t); }
    if (dup || (key >> depth)) bad = 1; else { sh.tmp[rank] = key; sh.slot[rank] = t; }
  }
  __syncthreads();
  if (bad) return false;
  bool first = false; u32 e = 0, s0 = 0xFFFF, s1 = 0xFFFF;
  if (t < cnt) {
    e = sh.tmp[t] & ~1u;
    first = t == 0 || (sh.tmp[t - 1] & ~1u) != e;
    if (first) {
      const bool has0 = sh.tmp[t] == e, has1 = has0 ? (t + 1 < cnt && sh.tmp[t + 1] == e + 1) : true;
      if (has0) s0 = sh.slot[t];
      if (has1) s1 = sh.slot[has0 ? t + 1 : t];
    }
  }
  u32 nidx; const u32 k0 = block_excl_sum(sh, first ? 1u : 0u, nidx);
  if (nidx != op.num_vecs) return false;
  if (first) { sh.idxA[k0] = e; sh.idxB[k0] = s0 | (s1 << 16); }
  __syncthreads();
  // per-position state (position = thread index, as the reference indexes `nodes` and `ptr` by position in the current level)
  const uint2 my_vec = t < nidx ? vec_index[op.idx_off + t] : make_uint2(0, 0);
  const u32 my_nodes = my_vec.x, my_cnt = my_vec.y;
  u32 ptr = 0;
  Digest* cd = sh.digA; Digest* nd = sh.digB; u32* ci = sh.idxA; u32* ni = sh.idxB;
  {
    Digest mine; u32 parent = 0;
    if (t < nidx) {
      const u32 ee = sh.idxA[t], sl = sh.idxB[t], a0 = sl & 0xFFFF, a1 = sl >> 16;
      Digest b0, b1;
      if (a0 != 0xFFFF && a1 != 0xFFFF) { b0 = sh.leaf[a0]; b1 = sh.leaf[a1]; ptr = 0; }
      else if (my_cnt == 0) { bad = 1; b0 = b1 = Digest{}; }
      else if (a0 != 0xFFFF) { b0 = sh.leaf[a0]; b1 = ld_digest(pb, my_nodes); ptr = 1; }
      else { b0 = ld_digest(pb, my_nodes); b1 = sh.leaf[a1]; ptr = 1; }
      mine = b3_merge(b0, b1); parent = ((1u << depth) + ee) >> 1;
    }
    __syncthreads();
    if (t < nidx) { cd[t] = mine; ci[t] = parent; }
    __syncthreads();
    if (bad) return false;
  }
  u32 n = nidx;
  for (u32 d = 1; d < depth; d++) {
    bool keep = false, pair = false; u32 node = 0;
    if (t < n) {
      node = ci[t];
      const bool second = (node & 1) && t > 0 && ci[t - 1] == node - 1;
      pair = !(node & 1) && t + 1 < n && ci[t + 1] == node + 1;
      keep = !second;
    }
    u32 nn; const u32 k = block_excl_sum(sh, keep ? 1u : 0u, nn);
    if (keep) {
      Digest sib;
      if (pair) sib = cd[t + 1];
      else if (ptr < my_cnt) { sib = ld_digest(pb, (size_t)my_nodes + 32 * (size_t)ptr); ptr++; }
      else { bad = 1; sib = Digest{}; }
      nd[k] = (node & 1) ? b3_merge(sib, cd[t]) : b3_merge(cd[t], sib);
      ni[k] = node >> 1;
    }
    __syncthreads();
    if (bad) return false;
    Digest* td = cd; cd = nd; nd = td; u32* ti = ci; ci = ni; ni = ti; n = nn;
  }
  if (n != 1 || ci[0] != 1) return false;
  root = cd[0];
  __syncthreads();
  return true;
}

template <int D> __device__ void verify_one(const VerifyRec& rec, const u8* __restrict__ pb, const uint2* __restrict__ vec_index, int* result, VShared& sh) {
  const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const u32 ln = rec.ln, lN = ln + 3, L = rec.num_layers;
  if (t == 0) { sh.status = 0; sh.npos = 0; }
  __syncthreads();

  // ---- (2) transcript up to the OOD point: coin seed, constraint composition coefficients, z ----
  Coin c;
  if (warp == 0) {
    c.seed = b3_hash_limbs<8 + XFG_NUM_PUB_INPUTS>(rec.seed_limbs); c.counter = 0;
    coin_reseed(c, rec.commitments[0]);
    bool ok = coin_draw_many<D>(c, XFG_NUM_TRANSITION + XFG_NUM_ASSERTIONS, sh.coef);
    coin_reseed(c, rec.commitments[1]);
    ok &= coin_draw_many<D>(c, 1, sh.z);
    if (!ok && lane == 0) vfail(sh, XFG_VERIFY_COIN);
  }
  __syncthreads();

  if (warp == 1 && lane == 0) {
    // ---- (3) OOD consistency: sum_k a_k r_k(T(z), T(zg)) (z - g^(n-1)) / (z^n - 1) + B0 / (z - 1) + B1 / (z - g^(n-1)) == H(z) ----
    const AirParams& air = rec.air;
    const Ext<D> z = lde2<D>(sh.z[0]);
    Ext<D> cur[XFG_TRACE_WIDTH], r[XFG_NUM_TRANSITION];
    for (int j = 0; j < XFG_TRACE_WIDTH; j++) cur[j] = lde2<D>(rec.ood_frame[2 * j]);
    const Ext<D> nxt4 = lde2<D>(rec.ood_frame[2 * 4 + 1]);
    // src/burn_mint_air.rs:356-377 over the extension field
    r[0] = sub_base(cur[0], XFG_STD_BURN) * sub_base(cur[0], gl_mul(XFG_STD_BURN, 1000));
    r[1] = cur[1] - cur[0];
    r[2] = sub_base(cur[2], air.txn);
    r[3] = sub_base(cur[3], air.rcpt);
    const Ext<D> dd = nxt4 - cur[4]; r[4] = dd * sub_base(dd, 1);
    r[5] = sub_base(cur[5], air.nullifier);
    r[6] = sub_base(cur[6], air.commitment);
    Ext<D> tsum, b0;
    for (int k = 0; k < XFG_NUM_TRANSITION; k++) tsum = tsum + lde2<D>(sh.coef[k]) * r[k];
    Ext<D> zn = z; for (u32 i = 0; i < ln; i++) zn = zn * zn;
    const Ext<D> zl = sub_base(z, air.g_last);
    Ext<D> res = tsum * zl * ext_inv(sub_base(zn, 1));
    for (int j = 0; j < XFG_TRACE_WIDTH; j++) b0 = b0 + lde2<D>(sh.coef[XFG_NUM_TRANSITION + j]) * sub_base(cur[j], air.assert0[j]);
    const Ext<D> b1 = lde2<D>(sh.coef[XFG_NUM_TRANSITION + XFG_TRACE_WIDTH]) * sub_base(cur[4], XFG_FINAL_STATE);
    res = res + b0 * ext_inv(sub_base(z, 1)) + b1 * ext_inv(zl);
    if (!ext_eq(res, lde2<D>(rec.hz))) vfail(sh, XFG_VERIFY_INCONSISTENT_OOD);
  }
  if (warp == 0) {
    // ---- (2, 4, 5) rest of the transcript: OOD frame and H(z) into the coin, DEEP coefficients, FRI alphas, proof of work, positions ----
    u64 limbs[2 * XFG_TRACE_WIDTH * D];
#pragma unroll
    for (int i = 0; i < 2 * XFG_TRACE_WIDTH; i++)
#pragma unroll
      for (int l = 0; l < D; l++) limbs[i * D + l] = rec.ood_frame[i][l];
    coin_reseed(c, b3_hash_limbs<2 * XFG_TRACE_WIDTH * D>(limbs));
    u64 hl[2] = {rec.hz[0], rec.hz[1]};
    coin_reseed(c, b3_hash_limbs<D>(hl));
    bool ok = coin_draw_many<D>(c, XFG_TRACE_WIDTH + 1, sh.dcoef);
    for (u32 l = 0; l < L; l++) { coin_reseed(c, rec.commitments[2 + l]); ok &= coin_draw_many<D>(c, 1, &sh.alphas[l]); }
    coin_reseed(c, rec.commitments[2 + L]);
    if (!ok && lane == 0) vfail(sh, XFG_VERIFY_COIN);
    const Digest pw = b3_merge_int(c.seed, rec.nonce);
    const u64 head = (u64)pw.w[0] | ((u64)pw.w[1] << 32), gmask = rec.grinding >= 64 ? ~0ull : ((1ull << rec.grinding) - 1);
    if ((head & gmask) != 0) { if (lane == 0) vfail(sh, XFG_VERIFY_POW_FAILED); }
    else {
      // draw_integers(q, N, nonce) -> sort -> dedup (A.5)
      const u64 mask = (1ull << lN) - 1; const u32 q = rec.num_queries;
      for (u32 i = lane; i < q; i += 32) { const Digest d = b3_merge_int(pw, (u64)i + 1); sh.tmp[i] = (u32)(((u64)d.w[0] | ((u64)d.w[1] << 32)) & mask); }
      __syncwarp();
      for (u32 i = lane; i < q; i += 32) {
        const u32 v = sh.tmp[i]; u32 rk = 0;
        for (u32 j = 0; j < q; j++) rk += (sh.tmp[j] < v) || (sh.tmp[j] == v && j < i);
        sh.slot[rk] = v;
      }
      __syncwarp();
      u32 base = 0;
      for (u32 i0 = 0; i0 < q; i0 += 32) {
        const u32 i = i0 + lane; const bool keep = i < q && (i == 0 || sh.slot[i - 1] != sh.slot[i]);
        const u32 m = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) sh.pos[base + __popc(m & ((1u << lane) - 1))] = sh.slot[i];
        base += __popc(m);
      }
      if (lane == 0) { sh.npos = base; if (base != rec.num_unique) vfail(sh, XFG_VERIFY_NUM_QUERIES_MISMATCH); }
    }
  }
  __syncthreads();
  if (sh.status) { if (t == 0) *result = sh.status; return; }
  const u32 npos = sh.npos;

  // ---- (6) trace / constraint openings ----
  // (values are read with their canonicity check, as read_elems does; the two value sections are checked before the first Merkle check)
  const size_t tv_off = rec.op[0].vals_off, cv_off = rec.op[1].vals_off;
  if (rec.op[0].vals_count != npos * XFG_TRACE_WIDTH || rec.op[1].vals_count != npos * D) { if (t == 0) *result = XFG_VERIFY_MALFORMED; return; }
  Digest root;
  u64 trow[XFG_TRACE_WIDTH], crow[D];
  if (t < npos) {
    bool canon = true;
    for (int j = 0; j < XFG_TRACE_WIDTH; j++) { trow[j] = ld_u64(pb, tv_off + 8 * ((size_t)t * XFG_TRACE_WIDTH + j)); canon &= trow[j] < GL_P; }
    for (int l = 0; l < D; l++) { crow[l] = ld_u64(pb, cv_off + 8 * ((size_t)t * D + l)); canon &= crow[l] < GL_P; }
    if (!canon) vfail(sh, XFG_VERIFY_MALFORMED);
    sh.leaf[t] = b3_hash_limbs<XFG_TRACE_WIDTH>(trow);
  }
  __syncthreads();
  if (sh.status) { if (t == 0) *result = sh.status; return; }
  if (!batch_root(sh, sh.pos, npos, lN, rec.op[0], pb, vec_index, root) || !dig_eq(root, rec.commitments[0])) { if (t == 0) *result = XFG_VERIFY_TRACE_QUERY_MISMATCH; return; }
  if (t < npos) sh.leaf[t] = b3_hash_limbs<D>(crow);
  __syncthreads();
  if (!batch_root(sh, sh.pos, npos, lN, rec.op[1], pb, vec_index, root) || !dig_eq(root, rec.commitments[1])) { if (t == 0) *result = XFG_VERIFY_CONSTRAINT_QUERY_MISMATCH; return; }

  // ---- (7) DEEP composition at the queried points ----
  const u64 g_N = gl_root_of_unity(lN);
  if (t < npos) {
    const Ext<D> z = lde2<D>(sh.z[0]), zg = mul_base(z, gl_root_of_unity(ln));
    const u64 x = gl_mul(XFG_GENERATOR, gl_pow(g_N, sh.pos[t]));
    const Ext<D> i1 = ext_inv(Ext<D>::from_base(x) - z), i2 = ext_inv(Ext<D>::from_base(x) - zg);
    Ext<D> acc;
    for (int j = 0; j < XFG_TRACE_WIDTH; j++) {
      const Ext<D> tv = Ext<D>::from_base(trow[j]);
      acc = acc + lde2<D>(sh.dcoef[j]) * ((tv - lde2<D>(rec.ood_frame[2 * j])) * i1 + (tv - lde2<D>(rec.ood_frame[2 * j + 1])) * i2);
    }
    Ext<D> cv; for (int l = 0; l < D; l++) cv.set_limb(l, crow[l]);
    acc = acc + lde2<D>(sh.dcoef[XFG_TRACE_WIDTH]) * (cv - lde2<D>(rec.hz)) * i1;
    sh.evals[t][0] = acc.limb(0); sh.evals[t][1] = D == 2 ? acc.limb(1) : 0;
  }
  __syncthreads();

  // ---- (8) FRI ----
  u32 cnt = npos, ldom = lN; u64 gen = g_N, max_deg_plus_1 = u64(1) << ln;
  u32* pos = sh.pos; u32* fpos = sh.fpos;
  u64 (*ev)[2] = sh.evals; u64 (*nx)[2] = sh.nxt;
  for (u32 l = 0; l < L; l++) {
    const u32 lrow = ldom - 3, rmask = (1u << lrow) - 1;
    // fold_positions: p mod (domain / 8), first occurrence kept, order preserved (A.10)
    bool keep = false; u32 v = 0;
    if (t < cnt) { v = pos[t] & rmask; keep = true; for (u32 j = 0; j < t; j++) if ((pos[j] & rmask) == v) { keep = false; break; } }
    u32 nf; const u32 k = block_excl_sum(sh, keep ? 1u : 0u, nf);
    if (keep) fpos[k] = v;
    __syncthreads();
    const VerifyOpening& op = rec.op[2 + l];
    if (op.vals_count != nf * 8 * D) { if (t == 0) *result = XFG_VERIFY_MALFORMED; return; }
    const size_t voff = op.vals_off;
    u64 row[8 * D];
    if (t < nf) {
      bool canon = true;
      for (int i = 0; i < 8 * D; i++) { row[i] = ld_u64(pb, voff + 8 * ((size_t)t * 8 * D + i)); canon &= row[i] < GL_P; }
      if (!canon) vfail(sh, XFG_VERIFY_MALFORMED);
      sh.leaf[t] = b3_hash_limbs<8 * D>(row);
    }
    __syncthreads();
    if (sh.status) { if (t == 0) *result = sh.status; return; }
    if (!batch_root(sh, fpos, nf, lrow, op, pb, vec_index, root) || !dig_eq(root, rec.commitments[2 + l])) { if (t == 0) *result = XFG_VERIFY_FRI_LAYER_COMMITMENT_MISMATCH; return; }
    // the queried row must contain the previous layer's value
    if (t < cnt) {
      const u32 pv = pos[t] & rmask; u32 idx = 0; while (idx < nf && fpos[idx] != pv) idx++;
      bool same = idx < nf;
      for (int ll = 0; ll < D; ll++) same = same && ld_u64(pb, voff + 8 * (((size_t)idx * 8 + (pos[t] >> lrow)) * D + ll)) == ev[t][ll];
      if (!same) vfail(sh, XFG_VERIFY_FRI_INVALID_LAYER_FOLDING);
    }
    if (max_deg_plus_1 % 8 != 0) vfail(sh, XFG_VERIFY_FRI_DEGREE_TRUNCATION);
    // fold: next[i] = P_i(alpha), P_i interpolating row i over x_i * w_8^j, x_i = 7 * gen^folded[i] (constant offset, A.10)
    if (t < nf) {
      Ext<D> rv[8];
      for (int j = 0; j < 8; j++) for (int ll = 0; ll < D; ll++) rv[j].set_limb(ll, row[j * D + ll]);
      const u64 xinv = gl_inv(gl_mul(XFG_GENERATOR, gl_pow(gen, fpos[t])));
      const Ext<D> w = fold8<D>(rv, rec.fc, mul_base(lde2<D>(sh.alphas[l]), xinv));
      nx[t][0] = w.limb(0); nx[t][1] = D == 2 ? w.limb(1) : 0;
    }
    __syncthreads();
    if (sh.status) { if (t == 0) *result = sh.status; return; }
    { u32* tp = pos; pos = fpos; fpos = tp; u64 (*te)[2] = ev; ev = nx; nx = te; }
    cnt = nf; ldom = lrow; max_deg_plus_1 /= 8;
    gen = gl_sqr(gl_sqr(gl_sqr(gen)));
  }
  // remainder: commitment, degree, evaluations at the last folded positions
  {
    u64* rl = reinterpret_cast<u64*>(sh.digA);    // MAX_REMAINDER * 2 words of scratch (the Merkle buffers are free now)
    for (u32 i = t; i < rec.rem_len * D; i += VT) rl[i] = rec.remainder[i / D][i % D];
    __syncthreads();
    if (t == 0) {
      const Digest rc = b3_hash_limbs_dyn(rl, rec.rem_len * D);
      if (!dig_eq(rc, rec.commitments[2 + L])) vfail(sh, XFG_VERIFY_FRI_REMAINDER_COMMITMENT_MISMATCH);
      else if (rec.rem_len > max_deg_plus_1) vfail(sh, XFG_VERIFY_FRI_REMAINDER_DEGREE_MISMATCH);
    }
    __syncthreads();
    if (sh.status) { if (t == 0) *result = sh.status; return; }
    if (t < cnt) {
      const u64 x = gl_mul(XFG_GENERATOR, gl_pow(gen, pos[t]));
      Ext<D> acc;
      for (int i = (int)rec.rem_len - 1; i >= 0; i--) { Ext<D> ci; for (int ll = 0; ll < D; ll++) ci.set_limb(ll, rl[i * D + ll]); acc = mul_base(acc, x) + ci; }
      if (!ext_eq(acc, lde2<D>(ev[t]))) vfail(sh, XFG_VERIFY_FRI_INVALID_REMAINDER_FOLDING);
    }
    __syncthreads();
  }
  if (t == 0) *result = sh.status;
}

__global__ void __launch_bounds__(VT) verify_kernel(const VerifyRec* __restrict__ recs, const u8* __restrict__ data, const uint2* __restrict__ vec_index, int* __restrict__ results) {
  __shared__ VShared sh;
  const VerifyRec& rec = recs[blockIdx.x];
  if (rec.host_status) { if (threadIdx.x == 0) results[blockIdx.x] = (int)rec.host_status; return; }
  if (rec.D == 1) verify_one<1>(rec, data + rec.base, vec_index, results + blockIdx.x, sh);
  else verify_one<2>(rec, data + rec.base, vec_index, results + blockIdx.x, sh);
}

void launch_verify(cudaStream_t st, const VerifyRec* recs, const u8* data, const uint2* vec_index, int* results, u32 count) {
  if (!count) return;
  verify_kernel<<<count, VT, 0, st>>>(recs, data, vec_index, results); XFG_LAUNCHED(1);
}

}  // namespace xfg
