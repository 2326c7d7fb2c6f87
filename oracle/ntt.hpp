// oracle/ntt.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Restates the results of winter-math 0.8.4 `fft::{interpolate_poly, interpolate_poly_with_offset,
// evaluate_poly_with_offset}` and `polynom::{eval, syn_div_in_place}`:
// natural-order in, natural-order out; evaluate_poly_with_offset(p, offset, b)[i] = p(offset * w_N^i), N = b*len(p).
#pragma once
#include <map>
#include "field.hpp"

namespace orc {

extern int g_threads;

inline std::vector<u64> power_series(u64 base, size_t n, u64 first = 1) {
  std::vector<u64> r(n); u64 x = first; for (size_t i = 0; i < n; i++) { r[i] = x; x = fmul(x, base); } return r;
}
inline unsigned ilog2(size_t n) { unsigned k = 0; while ((size_t(1) << k) < n) k++; return k; }

// per-stage twiddles, contiguous: stage with butterfly span `len` uses w_len^j for j < len/2 (stored at offset len/2)
inline const std::vector<u64>& twiddles(size_t n, bool inverse) {
  static std::map<std::pair<size_t, bool>, std::vector<u64>> cache;
  auto key = std::make_pair(n, inverse);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  std::vector<u64> t(n ? n : 1, 1);
  for (size_t len = 2; len <= n; len <<= 1) {
    u64 w = root_of_unity(ilog2(len)); if (inverse) w = finv(w);
    u64 x = 1; for (size_t j = 0; j < len / 2; j++) { t[len / 2 + j] = x; x = fmul(x, w); }
  }
  return cache[key] = std::move(t);
}

template <class E> inline void bit_reverse(E* a, size_t n) {
  unsigned k = ilog2(n);
  for (size_t i = 0, j = 0; i < n; i++) {
    if (i < j) std::swap(a[i], a[j]);
    size_t bit = n >> 1; while (bit && (j & bit)) { j ^= bit; bit >>= 1; } j |= bit;   // reversed-increment
  }
  (void)k;
}
// in-place NTT, natural order in and out; a[k] <- sum_j a[j] w^(jk)
template <class E> inline void ntt_core(E* a, size_t n, bool inverse) {
  if (n <= 1) return;
  const u64* tw = twiddles(n, inverse).data();
  bit_reverse(a, n);
  for (size_t len = 2; len <= n; len <<= 1) {
    size_t half = len / 2; const u64* t = tw + half;
    for (size_t i = 0; i < n; i += len)
      for (size_t j = 0; j < half; j++) {
        E u = a[i + j], v = a[i + j + half].mul_base(t[j]);
        a[i + j] = u + v; a[i + j + half] = u - v;
      }
  }
}
template <class E> inline void ntt(std::vector<E>& a) { ntt_core(a.data(), a.size(), false); }
// winter-math fft::interpolate_poly: evaluations over {w^i} -> coefficients
template <class E> inline void interpolate_poly(std::vector<E>& a) {
  ntt_core(a.data(), a.size(), true);
  u64 ninv = finv((u64)a.size() % P);
  for (auto& x : a) x = x.mul_base(ninv);
}
// winter-math fft::interpolate_poly_with_offset: evaluations over {offset * w^i} -> coefficients
template <class E> inline void interpolate_poly_with_offset(std::vector<E>& a, u64 offset) {
  interpolate_poly(a);
  u64 oi = finv(offset), x = 1;
  for (auto& c : a) { c = c.mul_base(x); x = fmul(x, oi); }
}
// winter-math fft::evaluate_poly_with_offset: result[i] = p(offset * w_N^i), N = blowup * len(p)
template <class E> inline std::vector<E> evaluate_poly_with_offset(const std::vector<E>& p, u64 offset, size_t blowup) {
  size_t n = p.size(), N = n * blowup;
  std::vector<E> out(N);
  u64 g = root_of_unity(ilog2(N));
  twiddles(n, false);
#pragma omp parallel for num_threads(g_threads) schedule(dynamic) if (g_threads > 1 && n >= 4096)
  for (size_t k = 0; k < blowup; k++) {
    std::vector<E> tmp(n);
    u64 s = fmul(offset, fpow(g, k)), x = 1;
    for (size_t j = 0; j < n; j++) { tmp[j] = p[j].mul_base(x); x = fmul(x, s); }
    ntt_core(tmp.data(), n, false);
    for (size_t m = 0; m < n; m++) out[m * blowup + k] = tmp[m];
  }
  return out;
}
// polynom::eval (Horner), coefficients of type C (F1 or E) evaluated at x in E
template <class E, class C> inline E eval_poly(const std::vector<C>& c, E x) {
  E r = E::zero();
  for (size_t i = c.size(); i-- > 0;) { r = r * x; for (int l = 0; l < C::DEG; l++) r.set_limb(l, fadd(r.limb(l), c[i].limb(l))); }
  return r;
}
// polynom::syn_div_in_place(p, 1, z): divide by (x - z), remainder dropped; result has the same length with top coeff 0
template <class E> inline void syn_div_in_place(std::vector<E>& p, E z) {
  E c = E::zero();
  for (size_t i = p.size(); i-- > 0;) { E t = p[i] + z * c; p[i] = c; c = t; }
}
// O(n^2) DFT used only to validate ntt_core in tests
template <class E> inline std::vector<E> naive_dft(const std::vector<E>& a, u64 w) {
  size_t n = a.size(); std::vector<E> r(n);
  for (size_t k = 0; k < n; k++) { E acc = E::zero(); u64 wk = fpow(w, k), x = 1; for (size_t j = 0; j < n; j++) { acc = acc + a[j].mul_base(x); x = fmul(x, wk); } r[k] = acc; }
  return r;
}

}  // namespace orc
