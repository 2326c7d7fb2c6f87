// field.cuh — Goldilocks (p = 2^64 - 2^32 + 1) and its quadratic extension F_p[x]/(x^2 - x + 2) for sm_100a.
//
// Replaces winter-math 0.8.4 `fields::f64::BaseElement` / `QuadExtension` on the device.
// Values are kept canonical (< p) in HBM so every buffer is bit-comparable with the reference's `as_int()` view;
// there is no Montgomery form: 2^64 = 2^32 - 1 and 2^96 = -1 (mod p) make the direct reduction cheaper on a
// 32-bit integer datapath than a Montgomery step.
#pragma once
#include <cstdint>
#include "../../include/xfg/spec.h"

#if defined(__CUDACC__)
#define XFG_HD __host__ __device__ __forceinline__
#define XFG_D __device__ __forceinline__
#else
#define XFG_HD inline
#define XFG_D inline
#endif

namespace xfg {

typedef uint64_t u64; typedef uint32_t u32; typedef uint8_t u8;
static constexpr u64 GL_P = XFG_P;
static constexpr u64 GL_EPS = 0xFFFFFFFFULL;   // 2^64 mod p

XFG_HD u64 gl_add(u64 a, u64 b) { u64 s = a + b; if (s < a) s += GL_EPS; else if (s >= GL_P) s -= GL_P; return s; }
XFG_HD u64 gl_sub(u64 a, u64 b) { u64 d = a - b; if (a < b) d += GL_P; return d; }
XFG_HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }
XFG_HD u64 gl_dbl(u64 a) { return gl_add(a, a); }

// 128 -> 64 bit reduction, canonical result
XFG_HD u64 gl_reduce128(u64 lo, u64 hi) {
  u64 hh = hi >> 32, hl = hi & GL_EPS;
  u64 t0 = lo - hh; if (lo < hh) t0 -= GL_EPS;
  u64 t1 = hl * GL_EPS;
  u64 r = t0 + t1; if (r < t1) r += GL_EPS;
  if (r >= GL_P) r -= GL_P;
  return r;
}
XFG_HD u64 gl_mul(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return gl_reduce128(a * b, __umul64hi(a, b));
#else
  unsigned __int128 x = (unsigned __int128)a * b; return gl_reduce128((u64)x, (u64)(x >> 64));
#endif
}
XFG_HD u64 gl_sqr(u64 a) { return gl_mul(a, a); }
XFG_HD u64 gl_pow(u64 b, u64 e) { u64 r = 1; while (e) { if (e & 1) r = gl_mul(r, b); b = gl_mul(b, b); e >>= 1; } return r; }
XFG_HD u64 gl_sqr_n(u64 a, int n) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1   // keep the 64 squarings of an inversion out of the instruction stream (the callers are I-cache bound otherwise)
#endif
  for (int i = 0; i < n; i++) a = gl_sqr(a);
  return a;
}
// a^(p-2), p - 2 = (2^31 - 1) * 2^33 + (2^32 - 1); 0 -> 0
XFG_HD u64 gl_inv(u64 x) {
  u64 t2 = gl_mul(gl_sqr(x), x), t3 = gl_mul(gl_sqr(t2), x);
  u64 t6 = gl_mul(gl_sqr_n(t3, 3), t3), t12 = gl_mul(gl_sqr_n(t6, 6), t6), t24 = gl_mul(gl_sqr_n(t12, 12), t12);
  u64 t30 = gl_mul(gl_sqr_n(t24, 6), t6), t31 = gl_mul(gl_sqr(t30), x), t32 = gl_mul(gl_sqr(t31), x);
  return gl_mul(gl_sqr_n(t31, 33), t32);
}
// primitive 2^k-th root of unity = G^(2^(32-k))  (A.1)
XFG_HD u64 gl_root_of_unity(unsigned k) { u64 r = XFG_TWO_ADIC_ROOT; for (unsigned i = k; i < XFG_TWO_ADICITY; i++) r = gl_sqr(r); return r; }

// ---- extension element, generic over degree D in {1, 2} (and 3, below): limb l of element i lives in array l (SoA in HBM) ----
template <int D> struct Ext;
template <> struct Ext<1> {
  u64 a0;
  XFG_HD Ext() : a0(0) {}
  XFG_HD explicit Ext(u64 x) : a0(x) {}
  XFG_HD Ext(u64 x, u64) : a0(x) {}
  XFG_HD static Ext from_base(u64 b) { return Ext(b); }
  XFG_HD u64 limb(int) const { return a0; }
  XFG_HD void set_limb(int, u64 v) { a0 = v; }
};
template <> struct Ext<2> {
  u64 a0, a1;
  XFG_HD Ext() : a0(0), a1(0) {}
  XFG_HD explicit Ext(u64 x) : a0(x), a1(0) {}
  XFG_HD Ext(u64 x, u64 y) : a0(x), a1(y) {}
  XFG_HD static Ext from_base(u64 b) { return Ext(b, 0); }
  XFG_HD u64 limb(int i) const { return i ? a1 : a0; }
  XFG_HD void set_limb(int i, u64 v) { if (i) a1 = v; else a0 = v; }
};
XFG_HD Ext<1> operator+(Ext<1> a, Ext<1> b) { return Ext<1>(gl_add(a.a0, b.a0)); }
XFG_HD Ext<1> operator-(Ext<1> a, Ext<1> b) { return Ext<1>(gl_sub(a.a0, b.a0)); }
XFG_HD Ext<1> operator*(Ext<1> a, Ext<1> b) { return Ext<1>(gl_mul(a.a0, b.a0)); }
XFG_HD Ext<1> mul_base(Ext<1> a, u64 b) { return Ext<1>(gl_mul(a.a0, b)); }
XFG_HD Ext<1> add_base(Ext<1> a, u64 b) { return Ext<1>(gl_add(a.a0, b)); }
XFG_HD Ext<1> ext_inv(Ext<1> a) { return Ext<1>(gl_inv(a.a0)); }
XFG_HD bool is_zero(Ext<1> a) { return a.a0 == 0; }
XFG_HD Ext<2> operator+(Ext<2> a, Ext<2> b) { return Ext<2>(gl_add(a.a0, b.a0), gl_add(a.a1, b.a1)); }
XFG_HD Ext<2> operator-(Ext<2> a, Ext<2> b) { return Ext<2>(gl_sub(a.a0, b.a0), gl_sub(a.a1, b.a1)); }
// (a0,a1)(b0,b1) = (a0b0 - 2 a1b1, (a0+a1)(b0+b1) - a0b0)   (A.1)
XFG_HD Ext<2> operator*(Ext<2> a, Ext<2> b) {
  u64 z = gl_mul(a.a0, b.a0), w = gl_mul(a.a1, b.a1);
  return Ext<2>(gl_sub(z, gl_dbl(w)), gl_sub(gl_mul(gl_add(a.a0, a.a1), gl_add(b.a0, b.a1)), z));
}
XFG_HD Ext<2> mul_base(Ext<2> a, u64 b) { return Ext<2>(gl_mul(a.a0, b), gl_mul(a.a1, b)); }
XFG_HD Ext<2> add_base(Ext<2> a, u64 b) { return Ext<2>(gl_add(a.a0, b), a.a1); }
// norm N(a) = a0^2 + a0a1 + 2a1^2 in F_p; a^-1 = (a0 + a1, -a1) / N(a)
XFG_HD u64 ext_norm(Ext<2> a) { return gl_add(gl_add(gl_sqr(a.a0), gl_mul(a.a0, a.a1)), gl_dbl(gl_sqr(a.a1))); }
XFG_HD u64 ext_norm(Ext<1> a) { return a.a0; }
XFG_HD Ext<2> ext_inv_with_norm_inv(Ext<2> a, u64 ninv) { return Ext<2>(gl_mul(gl_add(a.a0, a.a1), ninv), gl_mul(gl_neg(a.a1), ninv)); }
XFG_HD Ext<1> ext_inv_with_norm_inv(Ext<1>, u64 ninv) { return Ext<1>(ninv); }
XFG_HD Ext<2> ext_inv(Ext<2> a) { return ext_inv_with_norm_inv(a, gl_inv(ext_norm(a))); }
XFG_HD bool is_zero(Ext<2> a) { return (a.a0 | a.a1) == 0; }
// ---- cubic extension F_p[x]/(x^3 - x - 1) (winter-math 0.8.4 `impl ExtensibleField<3> for f64::BaseElement`; FieldExtension::Cubic of
// `ProofOptions`, src/burn_mint_prover.rs:44-49).  Used by the general-options pipeline (general_bodies.cuh); the tuned 8/8 kernels stay on degrees 1 and 2.
template <> struct Ext<3> {
  u64 a[3];
  XFG_HD Ext() : a{0, 0, 0} {}
  XFG_HD explicit Ext(u64 x) : a{x, 0, 0} {}
  XFG_HD Ext(u64 x, u64 y, u64 z) : a{x, y, z} {}
  XFG_HD static Ext from_base(u64 b) { return Ext(b); }
  XFG_HD u64 limb(int i) const { return a[i]; }
  XFG_HD void set_limb(int i, u64 v) { a[i] = v; }
};
XFG_HD Ext<3> operator+(Ext<3> a, Ext<3> b) { return Ext<3>(gl_add(a.a[0], b.a[0]), gl_add(a.a[1], b.a[1]), gl_add(a.a[2], b.a[2])); }
XFG_HD Ext<3> operator-(Ext<3> a, Ext<3> b) { return Ext<3>(gl_sub(a.a[0], b.a[0]), gl_sub(a.a[1], b.a[1]), gl_sub(a.a[2], b.a[2])); }
// schoolbook product c0..c4, then x^3 = x + 1 and x^4 = x^2 + x
XFG_HD Ext<3> operator*(Ext<3> a, Ext<3> b) {
  const u64 c0 = gl_mul(a.a[0], b.a[0]), c1 = gl_add(gl_mul(a.a[0], b.a[1]), gl_mul(a.a[1], b.a[0]));
  const u64 c2 = gl_add(gl_add(gl_mul(a.a[0], b.a[2]), gl_mul(a.a[1], b.a[1])), gl_mul(a.a[2], b.a[0]));
  const u64 c3 = gl_add(gl_mul(a.a[1], b.a[2]), gl_mul(a.a[2], b.a[1])), c4 = gl_mul(a.a[2], b.a[2]);
  return Ext<3>(gl_add(c0, c3), gl_add(gl_add(c1, c3), c4), gl_add(c2, c4));
}
XFG_HD Ext<3> mul_base(Ext<3> a, u64 b) { return Ext<3>(gl_mul(a.a[0], b), gl_mul(a.a[1], b), gl_mul(a.a[2], b)); }
XFG_HD Ext<3> add_base(Ext<3> a, u64 b) { return Ext<3>(gl_add(a.a[0], b), a.a[1], a.a[2]); }
XFG_HD bool is_zero(Ext<3> a) { return (a.a[0] | a.a[1] | a.a[2]) == 0; }
// a^-1: the columns a, a x, a x^2 are the matrix of "multiply by a" in the basis 1, x, x^2; the first column of its inverse (cofactors over the
// determinant = the norm of a) holds the coefficients of a^-1.   a x = (a2, a0 + a2, a1),   a x^2 = (a1, a1 + a2, a0 + a2)
XFG_HD Ext<3> ext_inv(Ext<3> e) {
  const u64 m00 = e.a[0], m10 = e.a[1], m20 = e.a[2];
  const u64 m01 = e.a[2], m11 = gl_add(e.a[0], e.a[2]), m21 = e.a[1];
  const u64 m02 = e.a[1], m12 = gl_add(e.a[1], e.a[2]), m22 = m11;
  const u64 k0 = gl_sub(gl_mul(m11, m22), gl_mul(m12, m21)), k1 = gl_sub(gl_mul(m10, m22), gl_mul(m12, m20)), k2 = gl_sub(gl_mul(m10, m21), gl_mul(m11, m20));
  const u64 det = gl_add(gl_sub(gl_mul(m00, k0), gl_mul(m01, k1)), gl_mul(m02, k2));
  const u64 di = gl_inv(det);
  return Ext<3>(gl_mul(k0, di), gl_mul(gl_neg(k1), di), gl_mul(k2, di));
}
// a^-1 = adj(a) / N(a) with N(a) in the base field: lets a caller batch the base-field inversions of several extension elements
//   degree 1: adj = 1, N = a;   degree 2: adj = (a0 + a1, -a1), N = a0^2 + a0 a1 + 2 a1^2;   degree 3: first column of the adjugate, N = det (see ext_inv)
XFG_HD u64 ext_norm_adj(Ext<1> a, Ext<1>& adj) { adj = Ext<1>(1); return a.a0; }
XFG_HD u64 ext_norm_adj(Ext<2> a, Ext<2>& adj) { adj = Ext<2>(gl_add(a.a0, a.a1), gl_neg(a.a1)); return ext_norm(a); }
XFG_HD u64 ext_norm_adj(Ext<3> e, Ext<3>& adj) {
  const u64 m00 = e.a[0], m10 = e.a[1], m20 = e.a[2];
  const u64 m01 = e.a[2], m11 = gl_add(e.a[0], e.a[2]), m21 = e.a[1];
  const u64 m02 = e.a[1], m12 = gl_add(e.a[1], e.a[2]), m22 = m11;
  const u64 k0 = gl_sub(gl_mul(m11, m22), gl_mul(m12, m21)), k1 = gl_sub(gl_mul(m10, m22), gl_mul(m12, m20)), k2 = gl_sub(gl_mul(m10, m21), gl_mul(m11, m20));
  adj = Ext<3>(k0, gl_neg(k1), k2);
  return gl_add(gl_sub(gl_mul(m00, k0), gl_mul(m01, k1)), gl_mul(m02, k2));
}
// the element `x` (the basis element of limb 1) times a: used to assemble H(z) from the limb polynomials of the composition column
template <int D> XFG_HD Ext<D> ext_mul_x(Ext<D> a);
template <> XFG_HD Ext<1> ext_mul_x<1>(Ext<1> a) { return a; }
template <> XFG_HD Ext<2> ext_mul_x<2>(Ext<2> a) { return Ext<2>(gl_neg(gl_dbl(a.a1)), gl_add(a.a0, a.a1)); }          // x^2 = x - 2
template <> XFG_HD Ext<3> ext_mul_x<3>(Ext<3> a) { return Ext<3>(a.a[2], gl_add(a.a[0], a.a[2]), a.a[1]); }           // x^3 = x + 1
template <int D> XFG_HD Ext<D> ext_pow(Ext<D> b, u64 e) { Ext<D> r(1); while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; } return r; }

// power of a fixed base through a two-level table: base^e = lo[e & 4095] * hi[e >> 12]
static constexpr int POW_LO_BITS = 12;
static constexpr u32 POW_LO = 1u << POW_LO_BITS;
struct PowTable { const u64* lo; const u64* hi; };
XFG_HD u64 pow_lookup(const PowTable& t, u64 e) {
  u64 l = t.lo[e & (POW_LO - 1)], h = e >> POW_LO_BITS;
  return h ? gl_mul(l, t.hi[h]) : l;
}

}  // namespace xfg
