// oracle/field.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product; see oracle/README.md.
//
// Restates winter-math 0.8.4 `fields::f64::BaseElement` (p = 2^64 - 2^32 + 1) and its `QuadExtension`
// (reference call sites src/burn_mint_air.rs:16-19).  Winterfell keeps elements in
// Montgomery form internally; every observable value (hashing, serialisation) is the canonical integer, so
// this restatement works on canonical integers throughout.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "../include/xfg/spec.h"

namespace orc {

using u8 = uint8_t; using u32 = uint32_t; using u64 = uint64_t; using u128 = unsigned __int128;
static constexpr u64 P = XFG_P;
static constexpr u64 EPS = 0xFFFFFFFFULL;  // 2^64 mod p = 2^32 - 1

inline u64 fadd(u64 a, u64 b) { u64 s = a + b; if (s < a || s >= P) s -= P; return s; }
inline u64 fsub(u64 a, u64 b) { return a >= b ? a - b : a + (P - b); }
inline u64 fneg(u64 a) { return a ? P - a : 0; }
// 128 -> 64 reduction: 2^64 = 2^32 - 1, 2^96 = -1 (mod p)
inline u64 freduce(u128 x) {
  u64 lo = (u64)x, hi = (u64)(x >> 64);
  u64 hh = hi >> 32, hl = hi & EPS;
  u64 t0 = lo - hh; if (lo < hh) t0 -= EPS;       // borrow: + p = - (2^32 - 1) mod 2^64
  u64 t1 = hl * EPS;
  u64 r = t0 + t1; if (r < t1) r += EPS;          // carry: 2^64 = 2^32 - 1
  if (r >= P) r -= P;
  return r;
}
inline u64 fmul(u64 a, u64 b) { return freduce((u128)a * b); }
inline u64 fmul_slow(u64 a, u64 b) { return (u64)(((u128)a * b) % P); }  // independent check of freduce
inline u64 fpow(u64 b, u64 e) { u64 r = 1; while (e) { if (e & 1) r = fmul(r, b); b = fmul(b, b); e >>= 1; } return r; }
inline u64 finv(u64 a) { return fpow(a, P - 2); }
// primitive 2^k-th root of unity: G^(2^(32-k))  (A.1; winter-math `get_root_of_unity`)
inline u64 root_of_unity(unsigned k) { u64 r = XFG_TWO_ADIC_ROOT; for (unsigned i = k; i < XFG_TWO_ADICITY; i++) r = fmul(r, r); return r; }

// ---- element types: F1 = base field, F2 = quadratic extension, F3 = cubic extension; same interface so the prover is generic ----
struct F1 {
  u64 v;
  static constexpr int DEG = 1;
  F1() : v(0) {}
  explicit F1(u64 x) : v(x) {}
  static F1 zero() { return F1(0); }
  static F1 one() { return F1(1); }
  static F1 from_base(u64 b) { return F1(b); }
  F1 operator+(F1 o) const { return F1(fadd(v, o.v)); }
  F1 operator-(F1 o) const { return F1(fsub(v, o.v)); }
  F1 operator*(F1 o) const { return F1(fmul(v, o.v)); }
  F1 operator-() const { return F1(fneg(v)); }
  F1 mul_base(u64 b) const { return F1(fmul(v, b)); }
  F1 inv() const { return F1(finv(v)); }
  bool operator==(F1 o) const { return v == o.v; }
  bool operator!=(F1 o) const { return v != o.v; }
  bool is_zero() const { return v == 0; }
  u64 limb(int) const { return v; }
  void set_limb(int, u64 x) { v = x; }
};

struct F2 {  // a0 + a1*x, x^2 = x - 2
  u64 a0, a1;
  static constexpr int DEG = 2;
  F2() : a0(0), a1(0) {}
  F2(u64 x0, u64 x1) : a0(x0), a1(x1) {}
  static F2 zero() { return F2(0, 0); }
  static F2 one() { return F2(1, 0); }
  static F2 from_base(u64 b) { return F2(b, 0); }
  F2 operator+(F2 o) const { return F2(fadd(a0, o.a0), fadd(a1, o.a1)); }
  F2 operator-(F2 o) const { return F2(fsub(a0, o.a0), fsub(a1, o.a1)); }
  F2 operator-() const { return F2(fneg(a0), fneg(a1)); }
  // winter-math ExtensibleField<2>::mul for f64: [a0b0 - 2 a1b1, (a0+a1)(b0+b1) - a0b0]   (A.1, D)
  F2 operator*(F2 o) const {
    u64 z = fmul(a0, o.a0), w = fmul(a1, o.a1);
    return F2(fsub(z, fadd(w, w)), fsub(fmul(fadd(a0, a1), fadd(o.a0, o.a1)), z));
  }
  F2 mul_base(u64 b) const { return F2(fmul(a0, b), fmul(a1, b)); }
  // conjugate (a0 + a1, -a1); norm a0^2 + a0a1 + 2a1^2   (A.1, D from QuadExtension::inv)
  F2 inv() const {
    u64 n = fadd(fadd(fmul(a0, a0), fmul(a0, a1)), fmul(2, fmul(a1, a1)));
    u64 ni = finv(n);
    return F2(fmul(fadd(a0, a1), ni), fmul(fneg(a1), ni));
  }
  bool operator==(F2 o) const { return a0 == o.a0 && a1 == o.a1; }
  bool operator!=(F2 o) const { return !(*this == o); }
  bool is_zero() const { return a0 == 0 && a1 == 0; }
  u64 limb(int i) const { return i ? a1 : a0; }
  void set_limb(int i, u64 x) { (i ? a1 : a0) = x; }
};

struct F3 {  // a0 + a1*x + a2*x^2, x^3 = x + 1  (winter-math 0.8.4 `impl ExtensibleField<3> for f64::BaseElement`: irreducible x^3 - x - 1;
             // pinned by the proofs the reference binary emits with FieldExtension::Cubic, tests/golden/reference_proofs_options.json)
  u64 a[3];
  static constexpr int DEG = 3;
  F3() : a{0, 0, 0} {}
  F3(u64 x0, u64 x1, u64 x2) : a{x0, x1, x2} {}
  static F3 zero() { return F3(); }
  static F3 one() { return F3(1, 0, 0); }
  static F3 from_base(u64 b) { return F3(b, 0, 0); }
  F3 operator+(F3 o) const { return F3(fadd(a[0], o.a[0]), fadd(a[1], o.a[1]), fadd(a[2], o.a[2])); }
  F3 operator-(F3 o) const { return F3(fsub(a[0], o.a[0]), fsub(a[1], o.a[1]), fsub(a[2], o.a[2])); }
  F3 operator-() const { return F3(fneg(a[0]), fneg(a[1]), fneg(a[2])); }
  // schoolbook product c0..c4, then x^3 = x + 1, x^4 = x^2 + x
  F3 operator*(F3 o) const {
    const u64* b = o.a;
    u64 c0 = fmul(a[0], b[0]), c1 = fadd(fmul(a[0], b[1]), fmul(a[1], b[0])), c2 = fadd(fadd(fmul(a[0], b[2]), fmul(a[1], b[1])), fmul(a[2], b[0]));
    u64 c3 = fadd(fmul(a[1], b[2]), fmul(a[2], b[1])), c4 = fmul(a[2], b[2]);
    return F3(fadd(c0, c3), fadd(fadd(c1, c3), c4), fadd(c2, c4));
  }
  F3 mul_base(u64 b) const { return F3(fmul(a[0], b), fmul(a[1], b), fmul(a[2], b)); }
  // inverse through the matrix of "multiply by this element" in the basis 1, x, x^2 (columns e * 1, e * x, e * x^2): the first column of
  // its inverse (Cramer's rule) holds the coefficients of e^-1.  The field is unique, so the value equals winter-math's Frobenius-based inv().
  F3 inv() const {
    const F3 c0 = *this, c1 = *this * F3(0, 1, 0), c2 = *this * F3(0, 0, 1);
    // M = [c0 c1 c2] (columns); solve M b = (1, 0, 0)
    auto det2 = [](u64 p, u64 q, u64 r, u64 s) { return fsub(fmul(p, s), fmul(q, r)); };
    const u64 m00 = c0.a[0], m10 = c0.a[1], m20 = c0.a[2], m01 = c1.a[0], m11 = c1.a[1], m21 = c1.a[2], m02 = c2.a[0], m12 = c2.a[1], m22 = c2.a[2];
    const u64 k0 = det2(m11, m12, m21, m22), k1 = det2(m10, m12, m20, m22), k2 = det2(m10, m11, m20, m21);
    const u64 det = fadd(fsub(fmul(m00, k0), fmul(m01, k1)), fmul(m02, k2));
    const u64 di = finv(det);
    return F3(fmul(k0, di), fmul(fneg(k1), di), fmul(k2, di));
  }
  bool operator==(F3 o) const { return a[0] == o.a[0] && a[1] == o.a[1] && a[2] == o.a[2]; }
  bool operator!=(F3 o) const { return !(*this == o); }
  bool is_zero() const { return a[0] == 0 && a[1] == 0 && a[2] == 0; }
  u64 limb(int i) const { return a[i]; }
  void set_limb(int i, u64 x) { a[i] = x; }
};

template <class E> inline E epow(E b, u64 e) { E r = E::one(); while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; } return r; }

// canonical little-endian serialisation, 8 bytes per base limb (A.1/A.6, D)
inline void put_u64(std::vector<u8>& out, u64 v) { for (int i = 0; i < 8; i++) out.push_back((u8)(v >> (8 * i))); }
inline u64 get_u64(const u8* p) { u64 v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | p[i]; return v; }
template <class E> inline void put_elem(std::vector<u8>& out, const E& e) { for (int i = 0; i < E::DEG; i++) put_u64(out, e.limb(i)); }
template <class E> inline void put_elems(std::vector<u8>& out, const std::vector<E>& es) { for (auto& e : es) put_elem(out, e); }

// winter-math `batch_inversion` (zeros stay zero)
template <class E> inline std::vector<E> batch_inverse(const std::vector<E>& v) {
  std::vector<E> r(v.size()); E acc = E::one();
  for (size_t i = 0; i < v.size(); i++) { r[i] = acc; if (!v[i].is_zero()) acc = acc * v[i]; }
  acc = acc.inv();
  for (size_t i = v.size(); i-- > 0;) { if (v[i].is_zero()) { r[i] = E::zero(); continue; } E t = r[i] * acc; acc = acc * v[i]; r[i] = t; }
  return r;
}

}  // namespace orc
