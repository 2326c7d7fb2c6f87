// b3_rolled.cuh — BLAKE3 for the latency-bound transcript code: ONE non-inlined, rolled compression per translation unit.
//
// The Fiat-Shamir steps (winter-crypto 0.8.3 `DefaultRandomCoin::{reseed, draw}`, `hash_elements` of the OOD frame / remainder)
// run once per proof on one warp, so every instruction they execute is a cold instruction-cache miss: with the unrolled compression of
// blake3.cuh (13 KB of SASS per inlined copy, several copies per kernel) code size, not arithmetic, set their latency.  b3r is one round in a
// 7-iteration loop plus the message permutation (~2.5 KB), shared by every caller of the unit.
#pragma once
#include "coin.cuh"

namespace xfg {

struct Msg { u32 w[16]; };

static __device__ __noinline__ Digest b3r(const Digest cv, Msg m, u32 block_len, u32 flags, u32 chunk_counter) {
  u32 s0 = cv.w[0], s1 = cv.w[1], s2 = cv.w[2], s3 = cv.w[3], s4 = cv.w[4], s5 = cv.w[5], s6 = cv.w[6], s7 = cv.w[7];
  u32 s8 = B3_IV0, s9 = B3_IV1, s10 = B3_IV2, s11 = B3_IV3, s12 = chunk_counter, s13 = 0, s14 = block_len, s15 = flags;
#pragma unroll 1
  for (int r = 0; r < 7; r++) {
    XFG_B3_ROUND(m.w[0], m.w[1], m.w[2], m.w[3], m.w[4], m.w[5], m.w[6], m.w[7], m.w[8], m.w[9], m.w[10], m.w[11], m.w[12], m.w[13], m.w[14], m.w[15])
    Msg t;   // message schedule of the next round: m'[i] = m[perm[i]]
    t.w[0] = m.w[2]; t.w[1] = m.w[6]; t.w[2] = m.w[3]; t.w[3] = m.w[10]; t.w[4] = m.w[7]; t.w[5] = m.w[0]; t.w[6] = m.w[4]; t.w[7] = m.w[13];
    t.w[8] = m.w[1]; t.w[9] = m.w[11]; t.w[10] = m.w[12]; t.w[11] = m.w[5]; t.w[12] = m.w[9]; t.w[13] = m.w[14]; t.w[14] = m.w[15]; t.w[15] = m.w[8];
    m = t;
  }
  Digest d;
  d.w[0] = s0 ^ s8; d.w[1] = s1 ^ s9; d.w[2] = s2 ^ s10; d.w[3] = s3 ^ s11; d.w[4] = s4 ^ s12; d.w[5] = s5 ^ s13; d.w[6] = s6 ^ s14; d.w[7] = s7 ^ s15;
  return d;
}
__device__ __forceinline__ Digest iv_digest() { Digest d; d.w[0] = B3_IV0; d.w[1] = B3_IV1; d.w[2] = B3_IV2; d.w[3] = B3_IV3; d.w[4] = B3_IV4; d.w[5] = B3_IV5; d.w[6] = B3_IV6; d.w[7] = B3_IV7; return d; }
__device__ __forceinline__ Digest r_merge(const Digest& l, const Digest& r) {
  Msg m;
#pragma unroll
  for (int i = 0; i < 8; i++) { m.w[i] = l.w[i]; m.w[8 + i] = r.w[i]; }
  return b3r(iv_digest(), m, 64, B3_SINGLE, 0);
}
__device__ __forceinline__ Digest r_merge_int(const Digest& s, u64 v) {
  Msg m;
#pragma unroll
  for (int i = 0; i < 8; i++) { m.w[i] = s.w[i]; m.w[8 + i] = 0; }
  m.w[8] = (u32)v; m.w[9] = (u32)(v >> 32);
  return b3r(iv_digest(), m, 40, B3_SINGLE, 0);
}
// one chunk (<= 128 limbs) of a hash_elements stream; `root` marks the only chunk of a short message
static __device__ Digest r_chunk(const u64* limbs, int nl, u32 chunk_counter, bool root) {
  Digest cv = iv_digest();
  const int nb = nl == 0 ? 1 : (nl + 7) / 8;
#pragma unroll 1
  for (int b = 0; b < nb; b++) {
    Msg m;
#pragma unroll
    for (int i = 0; i < 8; i++) { const int li = b * 8 + i; const u64 v = li < nl ? limbs[li] : 0; m.w[2 * i] = (u32)v; m.w[2 * i + 1] = (u32)(v >> 32); }
    const int rem = nl - b * 8; const u32 len = rem >= 8 ? 64 : (rem > 0 ? rem * 8 : 0);
    const u32 flags = (b == 0 ? XFG_B3_CHUNK_START : 0) | (b == nb - 1 ? (XFG_B3_CHUNK_END | (root ? XFG_B3_ROOT : 0)) : 0);
    cv = b3r(cv, m, len, flags, chunk_counter);
  }
  return cv;
}
__device__ __forceinline__ Digest r_parent(const Digest& l, const Digest& r, bool root) {
  Msg m;
#pragma unroll
  for (int i = 0; i < 8; i++) { m.w[i] = l.w[i]; m.w[8 + i] = r.w[i]; }
  return b3r(iv_digest(), m, 64, XFG_B3_PARENT | (root ? XFG_B3_ROOT : 0), 0);
}
// hash_elements of nl <= 512 limbs (BLAKE3 tree mode over up to 4 chunks), as b3_hash_limbs_dyn
static __device__ Digest r_hash_limbs(const u64* limbs, int nl) {
  const int chunks = nl <= 128 ? 1 : (nl + 127) / 128;
  if (chunks == 1) return r_chunk(limbs, nl, 0, true);
  Digest cv[4];
  for (int c = 0; c < chunks; c++) { int cl = nl - c * 128; if (cl > 128) cl = 128; cv[c] = r_chunk(limbs + c * 128, cl, (u32)c, false); }
  if (chunks == 2) return r_parent(cv[0], cv[1], true);
  const Digest l = r_parent(cv[0], cv[1], false);
  if (chunks == 3) return r_parent(l, cv[2], true);
  return r_parent(l, r_parent(cv[2], cv[3], false), true);
}
// DefaultRandomCoin::draw, `count` times (as coin_draw_many, on the rolled compression)
template <int D> static __device__ bool r_draw_many(Coin& c, u32 count, u64 (*out)[2]) {
  u32 got = 0;
  for (int round = 0; round < 40 && got < count; round++) {
    const Digest d = r_merge_int(c.seed, c.counter + 1 + lane_id());
    const u64 v0 = (u64)d.w[0] | ((u64)d.w[1] << 32), v1 = (u64)d.w[2] | ((u64)d.w[3] << 32);
    const bool valid = v0 < GL_P && (D == 1 || v1 < GL_P);
    const u32 mask = __ballot_sync(0xFFFFFFFFu, valid), rank = __popc(mask & ((1u << lane_id()) - 1)), need = count - got;
    if (valid && rank < need) { out[got + rank][0] = v0; out[got + rank][1] = D == 2 ? v1 : 0; }
    const u32 nvalid = __popc(mask);
    if (nvalid >= need) { c.counter += __fns(mask, 0, need) + 1; got = count; }
    else { c.counter += 32; got += nvalid; }
  }
  __syncwarp();
  return got == count;
}

// reseed(d): seed = BLAKE3(seed || d), counter = 0
__device__ __forceinline__ void r_reseed(Coin& c, const Digest& d) { c.seed = r_merge(c.seed, d); c.counter = 0; }

}  // namespace xfg
