// oracle/coin.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Restates winter-crypto 0.8.3 `DefaultRandomCoin<Blake3_256>` (bound at src/burn_mint_air.rs:484-485).
#pragma once
#include <stdexcept>
#include "hash.hpp"

namespace orc {

struct RandomCoin {
  Digest seed; u64 counter = 0;
  RandomCoin() : seed{} {}
  // new(seed elements): seed = hash_elements(elements), counter = 0
  explicit RandomCoin(const std::vector<F1>& elems) : seed(hash_elements(elems)), counter(0) {}
  void reseed(const Digest& d) { seed = merge(seed, d); counter = 0; }
  Digest next() { counter += 1; return merge_with_int(seed, counter); }
  // draw::<E>(): first 8*DEG bytes of next(); every limb must be a canonical element, else retry (<= 1000 tries)
  template <class E> E draw() {
    for (int t = 0; t < XFG_COIN_MAX_DRAWS; t++) {
      Digest d = next(); E e; bool ok = true;
      for (int l = 0; l < E::DEG; l++) { u64 v = get_u64(d.data() + 8 * l); if (v >= P) { ok = false; break; } e.set_limb(l, v); }
      if (ok) return e;
    }
    throw std::runtime_error("FailedToDrawFieldElement");
  }
  u32 check_leading_zeros(u64 nonce) const {
    Digest d = merge_with_int(seed, nonce); u64 head = get_u64(d.data());
    return head == 0 ? 64 : (u32)__builtin_ctzll(head);   // trailing_zeros of the LE head (A.5, D)
  }
  std::vector<size_t> draw_integers(size_t num_values, size_t domain_size, u64 nonce) {
    if (domain_size & (domain_size - 1)) throw std::runtime_error("domain size must be a power of two");
    if (num_values >= domain_size) throw std::runtime_error("number of values must be smaller than domain size");
    seed = merge_with_int(seed, nonce); counter = 0;
    u64 mask = (u64)domain_size - 1;
    std::vector<size_t> values;
    for (int t = 0; t < XFG_COIN_MAX_DRAWS; t++) {
      Digest d = next();
      values.push_back((size_t)(get_u64(d.data()) & mask));
      if (values.size() == num_values) break;
    }
    if (values.size() < num_values) throw std::runtime_error("FailedToDrawIntegers");
    return values;
  }
};

}  // namespace orc
