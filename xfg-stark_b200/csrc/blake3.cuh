// blake3.cuh — BLAKE3-256 compression for sm_100a, one compression per thread, state and message in registers.
//
// Replaces winter-crypto 0.8.3 `hashers::Blake3_256::{hash_elements, merge, merge_with_int}` over blake3 1.8.2
// (bound as HashFn at src/burn_mint_air.rs:483).  Bulk messages (rows, node pairs) are one or two 64-byte blocks of a single chunk; only the FRI remainder commitment can exceed
// one 1024-byte chunk (up to 4), handled by the runtime-length tree-mode variant used by the transcript kernels.
#pragma once
#include "field.cuh"

namespace xfg {

struct Digest { u32 w[8]; };

static constexpr u32 B3_IV0 = 0x6A09E667u, B3_IV1 = 0xBB67AE85u, B3_IV2 = 0x3C6EF372u, B3_IV3 = 0xA54FF53Au,
                     B3_IV4 = 0x510E527Fu, B3_IV5 = 0x9B05688Cu, B3_IV6 = 0x1F83D9ABu, B3_IV7 = 0x5BE0CD19u;

XFG_HD u32 b3_rotr(u32 x, int n) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(x, x, n);
#else
  return (x >> n) | (x << (32 - n));
#endif
}
// XFG_B3_IMAD: ptxas already issues the 2-input additions of G (c += d) on the FMA pipe as `IMAD.IADD x, y, 0x1, z` and keeps a + b + m as
// one ALU-pipe IADD3 (10 ALU + 2 FMA instructions per G; the ALU pipe is the bound of the hashing kernels - xfg_pipe_probe measures
// 18.5 T instr/s for ALU-only or IMAD-only code, 27.6 T for a 2+2 mix).  1: a + b is computed as IMAD a = b * one + a with an opaque multiplier (a __constant__ word), the
// message word is added by a 2-input addition that ptxas may place on either pipe (8-9 ALU + 4-6 FMA per G)
#ifndef XFG_B3_IMAD
#define XFG_B3_IMAD 1
#endif
#if defined(__CUDACC__)
static __constant__ u32 g_b3_one = 1;     // opaque multiplier: keeps ptxas from folding x*1+y back into an addition it would fuse into IADD3
#endif
XFG_HD u32 b3_add3_split(u32 a, u32 b, u32 m) {
#if defined(__CUDA_ARCH__)
  u32 t; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(b), "r"(g_b3_one), "r"(a)); return t + m;
#else
  return a + b + m;
#endif
}
#define XFG_B3_ADD_CD(c, d) c = c + d
#if XFG_B3_IMAD == 0
#define XFG_B3_ADD_ABM(a, b, m) a = a + b + (m)
#else
#define XFG_B3_ADD_ABM(a, b, m) a = b3_add3_split(a, b, (m))
#endif
#define XFG_B3_G(a, b, c, d, mx, my)                    \
  XFG_B3_ADD_ABM(a, b, mx); d = b3_rotr(d ^ a, 16);     \
  XFG_B3_ADD_CD(c, d);      b = b3_rotr(b ^ c, 12);     \
  XFG_B3_ADD_ABM(a, b, my); d = b3_rotr(d ^ a, 8);      \
  XFG_B3_ADD_CD(c, d);      b = b3_rotr(b ^ c, 7);
#define XFG_B3_ROUND(m0, m1, m2, m3, m4, m5, m6, m7, m8, m9, m10, m11, m12, m13, m14, m15) \
  XFG_B3_G(s0, s4, s8, s12, m0, m1) XFG_B3_G(s1, s5, s9, s13, m2, m3)                      \
  XFG_B3_G(s2, s6, s10, s14, m4, m5) XFG_B3_G(s3, s7, s11, s15, m6, m7)                    \
  XFG_B3_G(s0, s5, s10, s15, m8, m9) XFG_B3_G(s1, s6, s11, s12, m10, m11)                  \
  XFG_B3_G(s2, s7, s8, s13, m12, m13) XFG_B3_G(s3, s4, s9, s14, m14, m15)

// cv: chaining value (8 words), m: 16 message words; returns the first 8 output words (truncated compression)
XFG_HD void b3_compress(const u32 cv[8], const u32 m[16], u32 block_len, u32 flags, u32 out[8], u32 chunk_counter = 0) {
  u32 s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
  u32 s8 = B3_IV0, s9 = B3_IV1, s10 = B3_IV2, s11 = B3_IV3, s12 = chunk_counter, s13 = 0, s14 = block_len, s15 = flags;
  // message schedule: round r uses m[perm^r(i)], perm = {2,6,3,10,7,0,4,13,1,11,12,5,9,14,15,8}; written out so that
  // every index is a compile-time constant and the words stay in registers
  XFG_B3_ROUND(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14], m[15])
  XFG_B3_ROUND(m[2], m[6], m[3], m[10], m[7], m[0], m[4], m[13], m[1], m[11], m[12], m[5], m[9], m[14], m[15], m[8])
  XFG_B3_ROUND(m[3], m[4], m[10], m[12], m[13], m[2], m[7], m[14], m[6], m[5], m[9], m[0], m[11], m[15], m[8], m[1])
  XFG_B3_ROUND(m[10], m[7], m[12], m[9], m[14], m[3], m[13], m[15], m[4], m[0], m[11], m[2], m[5], m[8], m[1], m[6])
  XFG_B3_ROUND(m[12], m[13], m[9], m[11], m[15], m[10], m[14], m[8], m[7], m[2], m[5], m[3], m[0], m[1], m[6], m[4])
  XFG_B3_ROUND(m[9], m[14], m[11], m[5], m[8], m[12], m[15], m[1], m[13], m[3], m[0], m[10], m[2], m[6], m[4], m[7])
  XFG_B3_ROUND(m[11], m[15], m[5], m[0], m[1], m[9], m[8], m[6], m[14], m[10], m[2], m[12], m[3], m[4], m[7], m[13])
  out[0] = s0 ^ s8; out[1] = s1 ^ s9; out[2] = s2 ^ s10; out[3] = s3 ^ s11;
  out[4] = s4 ^ s12; out[5] = s5 ^ s13; out[6] = s6 ^ s14; out[7] = s7 ^ s15;
}
XFG_HD void b3_iv(u32 cv[8]) { cv[0] = B3_IV0; cv[1] = B3_IV1; cv[2] = B3_IV2; cv[3] = B3_IV3; cv[4] = B3_IV4; cv[5] = B3_IV5; cv[6] = B3_IV6; cv[7] = B3_IV7; }

static constexpr u32 B3_SINGLE = XFG_B3_CHUNK_START | XFG_B3_CHUNK_END | XFG_B3_ROOT;

// hash of NL canonical field limbs (8 bytes LE each), NL*8 <= 1024: `hash_elements` of a row (A.6)
template <int NL> XFG_HD Digest b3_hash_limbs(const u64* limbs) {
  static_assert(NL >= 1 && NL <= 128, "single chunk only");
  constexpr int NB = (NL + 7) / 8;
  u32 cv[8]; b3_iv(cv);
  Digest d;
#pragma unroll
  for (int b = 0; b < NB; b++) {
    u32 m[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int li = b * 8 + i;
      u64 v = li < NL ? limbs[li] : 0;
      m[2 * i] = (u32)v; m[2 * i + 1] = (u32)(v >> 32);
    }
    int rem = NL - b * 8; u32 len = rem >= 8 ? 64 : rem * 8;
    u32 flags = (b == 0 ? XFG_B3_CHUNK_START : 0) | (b == NB - 1 ? (XFG_B3_CHUNK_END | XFG_B3_ROOT) : 0);
    b3_compress(cv, m, len, flags, b == NB - 1 ? d.w : cv);
  }
  return d;
}
// `merge`: BLAKE3(left || right)
XFG_HD Digest b3_merge(const Digest& l, const Digest& r) {
  u32 cv[8]; b3_iv(cv); u32 m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { m[i] = l.w[i]; m[8 + i] = r.w[i]; }
  Digest d; b3_compress(cv, m, 64, B3_SINGLE, d.w); return d;
}
// `merge_with_int`: BLAKE3(seed || value LE)
XFG_HD Digest b3_merge_int(const Digest& s, u64 v) {
  u32 cv[8]; b3_iv(cv); u32 m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { m[i] = s.w[i]; m[8 + i] = 0; }
  m[8] = (u32)v; m[9] = (u32)(v >> 32);
  Digest d; b3_compress(cv, m, 40, B3_SINGLE, d.w); return d;
}
// runtime-length variant for the transcript (not on the bulk path): one chunk of up to 128 limbs
XFG_HD void b3_chunk_dyn(const u64* limbs, int nl, u32 chunk_counter, bool root, u32 out[8]) {
  u32 cv[8]; b3_iv(cv);
  int nb = nl == 0 ? 1 : (nl + 7) / 8;
  for (int b = 0; b < nb; b++) {
    u32 m[16];
    for (int i = 0; i < 8; i++) { int li = b * 8 + i; u64 v = li < nl ? limbs[li] : 0; m[2 * i] = (u32)v; m[2 * i + 1] = (u32)(v >> 32); }
    int rem = nl - b * 8; u32 len = rem >= 8 ? 64 : (rem > 0 ? rem * 8 : 0);
    u32 flags = (b == 0 ? XFG_B3_CHUNK_START : 0) | (b == nb - 1 ? (XFG_B3_CHUNK_END | (root ? XFG_B3_ROOT : 0)) : 0);
    b3_compress(cv, m, len, flags, b == nb - 1 ? out : cv, chunk_counter);
  }
}
XFG_HD void b3_parent(const u32 l[8], const u32 r[8], bool root, u32 out[8]) {
  u32 cv[8]; b3_iv(cv); u32 m[16];
  for (int i = 0; i < 8; i++) { m[i] = l[i]; m[8 + i] = r[i]; }
  b3_compress(cv, m, 64, XFG_B3_PARENT | (root ? XFG_B3_ROOT : 0), out);
}
// hash of nl limbs, nl <= 512 (4 chunks of 1024 bytes): BLAKE3 tree mode, left subtree = largest power of two of chunks
XFG_HD Digest b3_hash_limbs_dyn(const u64* limbs, int nl) {
  Digest d;
  const int chunks = nl <= 128 ? 1 : (nl + 127) / 128;
  if (chunks == 1) { b3_chunk_dyn(limbs, nl, 0, true, d.w); return d; }
  u32 cv[4][8];
  for (int c = 0; c < chunks; c++) { int cl = nl - c * 128; if (cl > 128) cl = 128; b3_chunk_dyn(limbs + c * 128, cl, (u32)c, false, cv[c]); }
  if (chunks == 2) { b3_parent(cv[0], cv[1], true, d.w); return d; }
  u32 l[8]; b3_parent(cv[0], cv[1], false, l);
  if (chunks == 3) { b3_parent(l, cv[2], true, d.w); return d; }
  u32 r[8]; b3_parent(cv[2], cv[3], false, r);
  b3_parent(l, r, true, d.w); return d;
}

#if defined(__CUDACC__)
XFG_D Digest load_digest(const Digest* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p); uint4 a = q[0], b = q[1];
  Digest d; d.w[0] = a.x; d.w[1] = a.y; d.w[2] = a.z; d.w[3] = a.w; d.w[4] = b.x; d.w[5] = b.y; d.w[6] = b.z; d.w[7] = b.w; return d;
}
XFG_D void store_digest(Digest* p, const Digest& d) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(d.w[0], d.w[1], d.w[2], d.w[3]); q[1] = make_uint4(d.w[4], d.w[5], d.w[6], d.w[7]);
}
#endif

}  // namespace xfg
