// field_weak.cuh — "weak" Goldilocks arithmetic for the NTT inner loops (device only).
//
// A weak value is any u64 standing for its residue mod p = 2^64 - 2^32 + 1 (so p..2^64-1 alias 0..2^32-2); a canonical
// value is < p.  Butterflies keep their running values weak and only canonicalise the twiddled operand, which removes most
// compare/select work from the 32-bit datapath.  Every routine states which operands may be weak; the fix-ups rely on
//   2^64 = 2^32 - 1 (EPS),  2^96 = -1,  2^192 = 1  (mod p).
// A carry out of bit 64 is repaid by adding EPS, a borrow by subtracting EPS; each helper documents why a single fix suffices.
// Carry flags: ptxas implements `subc` with the hardware convention (carry-in = NOT borrow), so the borrow mask of a SUB chain
// is `subc m,0,0`, but the carry mask of an ADD chain must be built as `addc c,0,0; neg` (checked in SASS).
// These replace the inner arithmetic of winter-math 0.8.4 `fft::fft_inputs` butterflies.
#pragma once
#include "field.cuh"

namespace xfg {

__device__ __forceinline__ u64 w_pack(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

// XFG_WEAK_MADFIX (A/B switch, default 0): the "+ EPS when bit 64 was carried out" fix-up as one `mad.wide  r = c * EPS + r` (c = the carry)
// instead of mask-select + two-word addition, i.e. 2-3 ALU-pipe instructions traded for FMA-pipe work.  MEASURED SLOWER on B200 (2^20 rows,
// quadratic): ntt.lde_trace 1.154 -> 1.254 ms, deep 0.631 -> 0.649, constraints 0.195 -> 0.199 - ptxas expands the multiplication by 2^32 - 1 into
// IMAD.HI + IMAD + extra moves (IMAD.HI 86 -> 311 per NTT kernel) and the heavy half of the FMA pipe (IMAD.WIDE / IMAD.HI, 7.6 T instr/s against
// 18.5 T for 32-bit IMAD) is already as loaded as the ALU pipe.  Kept for the record; the select form stays the product path.
#ifndef XFG_WEAK_MADFIX
#define XFG_WEAK_MADFIX 0
#endif
// r + c * EPS for c in {0, 1} (mod 2^64)
#if XFG_WEAK_MADFIX >= 2
// 2, 3, 4 (round 2): the multiplier 2^32 - 1 comes from __constant__ memory, so ptxas cannot expand the product into IMAD.HI + IMAD + moves (one
// IMAD.WIDE per fix-up), and the carry is materialised on the FMA pipe (madc.lo = IMAD.X).  3: only the butterfly additions w_add_c use the product
// form; 4: w_add_c and w_canon.  Static SASS of ntt_pass_r16<.., 10> (tools/sass_pipes.py): ALU-pipe instructions 3229 -> 3148 (3) / 3034 (4), IMAD.WIDE
// + IMAD.HI 498 -> 578 / 692.  MEASURED SLOWER all the same (2^20 rows, quadratic; ntt.lde_trace 1.123 ms -> 1.135 (3) / 1.168 (4), whole proof 3.977 ->
// 3.999 / 4.061): an IMAD.WIDE costs the kernel more than the two ALU-pipe instructions it replaces.
static __constant__ u32 g_w_eps2 = 0xFFFFFFFFu;
__device__ __forceinline__ u64 w_fix_eps(u32 c, u64 r) { u64 o; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(o) : "r"(c), "r"(g_w_eps2), "l"(r)); return o; }
#else
__device__ __forceinline__ u64 w_fix_eps(u32 c, u64 r) { u64 o; asm("mad.wide.u32 %0, %1, 0xffffffff, %2;" : "=l"(o) : "r"(c), "l"(r)); return o; }
#endif

// (Round 2, not kept: the fix-up as a PREDICATED mad.wide on the carry predicate with multiplicands from __constant__ / device memory.  ptxas splits
// the mad, hoists the loop-invariant product 1 * EPS into a register pair and if-converts the guarded addition back to IADD3 + IMAD.X + 2 SEL, i.e.
// more ALU-pipe work than the select form: 3229 -> 3932 ALU instructions in ntt_pass_r16<.., 10>; see tools/sass_pipes.py.)
// weak -> canonical.  x >= p  <=>  x + EPS carries out of bit 64, and then x - p = x + EPS (mod 2^64)
__device__ __forceinline__ u64 w_canon(u64 x) {
#if XFG_WEAK_MADFIX == 4
  u32 c;       // 4 = 3 + w_canon: carry of x + EPS materialised on the FMA pipe (madc.lo), then x + c * EPS as one IMAD.WIDE
  asm("{\n\t .reg .u32 t0, t1;\n\t add.cc.u32 t0, %1, 0xffffffff;\n\t addc.cc.u32 t1, %2, 0;\n\t madc.lo.u32 %0, %3, %3, 0;\n\t}"
      : "=r"(c) : "r"((u32)x), "r"((u32)(x >> 32)), "r"(0u));
  return w_fix_eps(c, x);
#elif XFG_WEAK_MADFIX == 1 || XFG_WEAK_MADFIX == 2
  u32 c;
  asm("{\n\t .reg .u32 t0, t1;\n\t add.cc.u32 t0, %1, 0xffffffff;\n\t addc.cc.u32 t1, %2, 0;\n\t addc.u32 %0, 0, 0;\n\t}"
      : "=r"(c) : "r"((u32)x), "r"((u32)(x >> 32)));
  return w_fix_eps(c, x);
#else
  u32 r0, r1;
  asm("{\n\t .reg .pred p; .reg .u32 c, t0, t1;\n\t add.cc.u32 t0, %2, 0xffffffff;\n\t addc.cc.u32 t1, %3, 0;\n\t addc.u32 c, 0, 0;\n\t setp.ne.u32 p, c, 0;\n\t"
      "selp.b32 %0, t0, %2, p;\n\t selp.b32 %1, t1, %3, p;\n\t}"
      : "=r"(r0), "=r"(r1) : "r"((u32)x), "r"((u32)(x >> 32)));
  return w_pack(r0, r1);
#endif
}

// a (weak) + b (canonical) -> weak.  a + b < 2^64 + p, so after a carry the wrapped sum is < p and adding EPS cannot carry again.
__device__ __forceinline__ u64 w_add_c(u64 a, u64 b) {
#if XFG_WEAK_MADFIX
  u32 s0, s1, c;
  asm("{\n\t add.cc.u32 %0, %3, %5;\n\t addc.cc.u32 %1, %4, %6;\n\t madc.lo.u32 %2, %7, %7, 0;\n\t}"
      : "=&r"(s0), "=&r"(s1), "=&r"(c) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)), "r"(0u));
  return w_fix_eps(c, w_pack(s0, s1));
#else
  // The carry mask is built as setp/selp so that ptxas keeps it one SEL on the carry predicate (5 instructions in all).
  u32 s0, s1;
  asm("{\n\t .reg .pred p; .reg .u32 c, m;\n\t add.cc.u32 %0, %2, %4;\n\t addc.cc.u32 %1, %3, %5;\n\t addc.u32 c, 0, 0;\n\t setp.ne.u32 p, c, 0;\n\t"
      "selp.b32 m, 0xffffffff, 0, p;\n\t add.cc.u32 %0, %0, m;\n\t addc.u32 %1, %1, 0;\n\t}"
      : "=&r"(s0), "=&r"(s1) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
  return w_pack(s0, s1);
#endif
}
// a (weak) - b (canonical) -> weak.  After a borrow the wrapped difference is >= 2^64 - (p-1) >= EPS, so subtracting EPS cannot borrow again.
__device__ __forceinline__ u64 w_sub_c(u64 a, u64 b) {
  u32 s0, s1, m;
  asm("{\n\t sub.cc.u32 %0, %3, %5;\n\t subc.cc.u32 %1, %4, %6;\n\t subc.u32 %2, 0, 0;\n\t sub.cc.u32 %0, %0, %2;\n\t subc.u32 %1, %1, 0;\n\t}"
      : "=&r"(s0), "=&r"(s1), "=&r"(m) : "r"((u32)a), "r"((u32)(a >> 32)), "r"((u32)b), "r"((u32)(b >> 32)));
  return w_pack(s0, s1);
}
// a * EPS = a * 2^64 (mod p) for a < 2^32: one IMAD.WIDE on the FMA pipe; the result is canonical (<= 2^64 - 2^33 + 1 < p)
__device__ __forceinline__ u64 w_mul_eps(u32 a) { u64 r; asm("mul.wide.u32 %0, %1, 0xffffffff;" : "=l"(r) : "r"(a)); return r; }
// a * EPS + (b1:b0) for a < 2^32 and any u64 (b1:b0) -> weak: IMAD.WIDE with the 64-bit addend, then the carry fix.  The high word of
// a * EPS is at most 2^32 - 2, so bit 64 was carried out exactly when the sum's high word is below b1 (one 32-bit compare); the
// wrapped sum is then <= 2^64 - 2^33, so + EPS cannot carry again.
__device__ __forceinline__ u64 w_eps_madd(u32 a, u32 b0, u32 b1) {
#if XFG_WEAK_MADFIX == 1 || XFG_WEAK_MADFIX == 2
  u64 r; u32 c;
  asm("{\n\t .reg .u64 b; .reg .pred p; .reg .u32 h;\n\t mov.b64 b, {%3, %4};\n\t mad.wide.u32 %0, %2, 0xffffffff, b;\n\t mov.b64 {_, h}, %0;\n\t"
      "setp.lt.u32 p, h, %4;\n\t selp.u32 %1, 1, 0, p;\n\t}"
      : "=&l"(r), "=&r"(c) : "r"(a), "r"(b0), "r"(b1));
  return w_fix_eps(c, r);
#else
  u32 r0, r1;
  asm("{\n\t .reg .u64 r, b; .reg .pred p; .reg .u32 m;\n\t mov.b64 b, {%3, %4};\n\t mad.wide.u32 r, %2, 0xffffffff, b;\n\t mov.b64 {%0, %1}, r;\n\t"
      "setp.lt.u32 p, %1, %4;\n\t selp.b32 m, 0xffffffff, 0, p;\n\t add.cc.u32 %0, %0, m;\n\t addc.u32 %1, %1, 0;\n\t}"
      : "=&r"(r0), "=&r"(r1) : "r"(a), "r"(b0), "r"(b1));
  return w_pack(r0, r1);
#endif
}
// r (weak) + c * 2^32, c < 2^32 -> weak
__device__ __forceinline__ u64 w_add_hi32(u64 r, u32 c) { return w_add_c(r, w_pack(0u, c)); }   // c * 2^32 <= 2^64 - 2^32 < p: canonical
// r (weak) - c * 2^32, c < 2^32 -> weak
__device__ __forceinline__ u64 w_sub_hi32(u64 r, u32 c) { return w_sub_c(r, w_pack(0u, c)); }
// r (weak) - c, c < 2^32 -> weak
__device__ __forceinline__ u64 w_sub32(u64 r, u32 c) { return w_sub_c(r, (u64)c); }

// 128-bit (lo, hi) -> weak residue:  (lo - hi_hi) + hi_lo * EPS.  lo is any u64 (weak), hi_hi < 2^32 is canonical.
__device__ __forceinline__ u64 w_reduce128(u64 lo, u64 hi) { const u64 t = w_sub_c(lo, hi >> 32); return w_eps_madd((u32)hi, (u32)t, (u32)(t >> 32)); }
// a * b for any u64 a, b (weak operands allowed) -> weak
__device__ __forceinline__ u64 w_mul(u64 a, u64 b) {
  const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  const u64 ll = (u64)a0 * b0;
  const u64 t1 = (u64)a0 * b1 + (ll >> 32);          // none of these sums can overflow 64 bits
  const u64 t2 = (u64)a1 * b0 + (u32)t1;
  const u64 hi = (u64)a1 * b1 + (t1 >> 32) + (t2 >> 32);
  const u64 lo = (t2 << 32) | (u32)ll;
  return w_reduce128(lo, hi);
}

// x^(p-2) with weak intermediate products (same addition chain as gl_inv; one canonicalisation at the end): the batch inversions of the
// constraint and DEEP kernels spend a fifth to a third of their instructions here.  x canonical or weak; 0 -> 0.
__device__ __forceinline__ u64 w_sqr_n(u64 a, int n) {
#pragma unroll 1
  for (int i = 0; i < n; i++) a = w_mul(a, a);
  return a;
}
__device__ __forceinline__ u64 w_inv(u64 x) {
  const u64 t2 = w_mul(w_mul(x, x), x), t3 = w_mul(w_mul(t2, t2), x);
  const u64 t6 = w_mul(w_sqr_n(t3, 3), t3), t12 = w_mul(w_sqr_n(t6, 6), t6), t24 = w_mul(w_sqr_n(t12, 12), t12);
  const u64 t30 = w_mul(w_sqr_n(t24, 6), t6), t31 = w_mul(w_mul(t30, t30), x), t32 = w_mul(w_mul(t31, t31), x);
  return w_canon(w_mul(w_sqr_n(t31, 33), t32));
}

// Block-wide batch inversion (Montgomery's trick across the CTA): thread i hands in a non-zero value (weak allowed; 1 for a thread without work) and
// receives its inverse (weak).  Warp 0 multiplies the block's THREADS values up (lane l owns values l, l + 32, ...), runs the ONE x^(p-2) chain and
// unwinds; the other warps wait at the barrier.  The chain is 72 dependent multiplications however many lanes run it, so one chain per block issues
// 32 / THREADS of the warp-instructions that one chain per warp does.  buf: THREADS words of shared memory.  Every thread of the block must call it.
template <int THREADS> __device__ __forceinline__ u64 cta_inv(u64 v, u64* buf) {
  constexpr int G = THREADS / 32;
  static_assert(THREADS % 32 == 0 && G >= 1, "whole warps");
  const u32 tid = threadIdx.x;
  buf[tid] = v;
  __syncthreads();
  if (tid < 32) {
    u64 pre[G]; u64 acc = buf[tid];
#pragma unroll
    for (int g = 1; g < G; g++) { pre[g] = acc; acc = w_mul(acc, buf[g * 32 + tid]); }
    acc = w_inv(acc);
#pragma unroll
    for (int g = G - 1; g >= 1; g--) { const u64 t = w_mul(pre[g], acc); acc = w_mul(acc, buf[g * 32 + tid]); buf[g * 32 + tid] = t; }
    buf[tid] = acc;
  }
  __syncthreads();
  return buf[tid];
}

// two-level power table lookup with a weak result (see pow_lookup in field.cuh)
__device__ __forceinline__ u64 w_pow_lookup(const PowTable& t, u64 e) {
  const u64 l = t.lo[e & (POW_LO - 1)], h = e >> POW_LO_BITS;
  return h ? w_mul(l, t.hi[h]) : l;
}

// Dot-product accumulator: sum of up to 2^32 full 128-bit products kept un-reduced in 160 bits (lo, hi, top) and reduced once:
//   lo + hi*2^64 + top*2^128 = reduce128(lo, hi) - top*2^32   (2^128 = -2^32 mod p)
// One term costs the 4 IMAD.WIDE of the product + 5 carry-chain additions instead of a full multiply-reduce-add (~46 instr).
struct DotAcc {
  u64 lo, hi; u32 top;
  __device__ __forceinline__ DotAcc() : lo(0), hi(0), top(0) {}
  __device__ __forceinline__ void fma(u64 a, u64 b) {      // operands may be weak
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    const u64 ll = (u64)a0 * b0;
    const u64 t1 = (u64)a0 * b1 + (ll >> 32);
    const u64 t2 = (u64)a1 * b0 + (u32)t1;
    const u64 ph = (u64)a1 * b1 + (t1 >> 32) + (t2 >> 32);
    const u64 pl = (t2 << 32) | (u32)ll;
    asm("{\n\t add.cc.u64 %0, %0, %3;\n\t addc.cc.u64 %1, %1, %4;\n\t addc.u32 %2, %2, 0;\n\t}" : "+l"(lo), "+l"(hi), "+r"(top) : "l"(pl), "l"(ph));
  }
  __device__ __forceinline__ u64 result() const { return w_canon(w_sub_hi32(w_reduce128(lo, hi), top)); }   // canonical
};

// x * 2^S for a compile-time S in [0, 96), x weak -> weak.  With S = 32q + t and y = x << t = y2*2^64 + y1*2^32 + y0 (y2 < 2^t):
//   q = 0:  (y1:y0) + y2*EPS                            (2^64 = EPS)
//   q = 1:  (y0:0)  + y1*EPS - y2                       (2^96 = -1)
//   q = 2:  y0*EPS  - (y2:y1)                           (2^96 = -1, 2^128 = -2^32)
// The EPS products are single IMAD.WIDE instructions with the other term as the 64-bit addend (w_eps_madd; FMA pipe, which these
// kernels leave mostly idle): 8 / 13 / 9 instructions for q = 0 / 1 / 2.
template <int S> __device__ __forceinline__ u64 w_mul_pow2(u64 x) {
  static_assert(S >= 0 && S < 96, "shift out of range");
  if (S == 0) return x;
  constexpr int q = S / 32, t = S % 32;
  const u32 x0 = (u32)x, x1 = (u32)(x >> 32);
  const u32 y0 = t ? (x0 << t) : x0, y1 = t ? __funnelshift_l(x0, x1, t) : x1, y2 = t ? (x1 >> (32 - t)) : 0u;
  if (q == 0) return w_eps_madd(y2, y0, y1);
  if (q == 1) { const u64 r = w_eps_madd(y1, 0u, y0); return t ? w_sub32(r, y2) : r; }
  return w_sub_c(w_mul_eps(y0), w_pack(y1, y2));        // (y2:y1) < 2^63: canonical
}

// -x for a weak x -> weak:  ~x = 2^64 - 1 - x stands for EPS - 1 - x, so -x = ~x - (EPS - 1)
__device__ __forceinline__ u64 w_neg(u64 x) { return w_sub_c(~x, GL_EPS - 1); }
// x * 2^E for a compile-time E in [0, 192) (2 has order 192; 2^96 = -1), x weak -> weak
template <int E> __device__ __forceinline__ u64 w_mul_pow2_any(u64 x) {
  static_assert(E >= 0 && E < 192, "exponent out of range");
  if (E < 96) return w_mul_pow2<(E < 96 ? E : 0)>(x);
  return w_neg(w_mul_pow2<(E >= 96 ? E - 96 : 0)>(x));
}

}  // namespace xfg
