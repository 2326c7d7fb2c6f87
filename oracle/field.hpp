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

// ---- element types: F1 = base field, F2 = quadratic extension; same interface so the prover is generic ----
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
