// oracle/hash.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// BLAKE3-256 (default hash mode, full tree mode for inputs > 1024 B) restating the published BLAKE3 spec as used
// by blake3 1.8.2 behind winter-crypto 0.8.3 `Blake3_256` (binding at src/burn_mint_air.rs:483),
// and Keccak-256 (original Keccak padding 0x01, as sha3 0.10 `Keccak256`; call sites src/burn_mint_air.rs:124-202,
// src/burn_mint_prover.rs:211-221).  Pinned in tests against Python `blake3` and the reference KAT src/lib.rs:141-148.
#pragma once
#include <array>
#include "field.hpp"

namespace orc {

using Digest = std::array<u8, 32>;

namespace b3 {
static constexpr u32 IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};
static constexpr int PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
inline u32 rotr(u32 x, int n) { return (x >> n) | (x << (32 - n)); }
inline void g(u32* s, int a, int b, int c, int d, u32 mx, u32 my) {
  s[a] = s[a] + s[b] + mx; s[d] = rotr(s[d] ^ s[a], 16);
  s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 12);
  s[a] = s[a] + s[b] + my; s[d] = rotr(s[d] ^ s[a], 8);
  s[c] = s[c] + s[d];      s[b] = rotr(s[b] ^ s[c], 7);
}
// full compression; out[0..8) is the chaining value / first 32 output bytes
inline void compress(const u32 cv[8], const u32 block[16], u64 counter, u32 block_len, u32 flags, u32 out[8]) {
  u32 s[16] = {cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], cv[6], cv[7], IV[0], IV[1], IV[2], IV[3],
               (u32)counter, (u32)(counter >> 32), block_len, flags};
  u32 m[16]; for (int i = 0; i < 16; i++) m[i] = block[i];
  for (int r = 0; r < 7; r++) {
    g(s, 0, 4, 8, 12, m[0], m[1]);  g(s, 1, 5, 9, 13, m[2], m[3]);
    g(s, 2, 6, 10, 14, m[4], m[5]); g(s, 3, 7, 11, 15, m[6], m[7]);
    g(s, 0, 5, 10, 15, m[8], m[9]);   g(s, 1, 6, 11, 12, m[10], m[11]);
    g(s, 2, 7, 8, 13, m[12], m[13]);  g(s, 3, 4, 9, 14, m[14], m[15]);
    if (r < 6) { u32 t[16]; for (int i = 0; i < 16; i++) t[i] = m[PERM[i]]; for (int i = 0; i < 16; i++) m[i] = t[i]; }
  }
  for (int i = 0; i < 8; i++) out[i] = s[i] ^ s[i + 8];
}
inline void load_block(const u8* p, size_t len, u32 w[16]) {
  u8 buf[64] = {0}; std::memcpy(buf, p, len);
  for (int i = 0; i < 16; i++) w[i] = (u32)buf[4 * i] | ((u32)buf[4 * i + 1] << 8) | ((u32)buf[4 * i + 2] << 16) | ((u32)buf[4 * i + 3] << 24);
}
// one chunk (<= 1024 bytes): returns cv; if `root`, the last block carries ROOT
inline void chunk_cv(const u8* p, size_t len, u64 chunk_counter, bool root, u32 out[8]) {
  u32 cv[8]; for (int i = 0; i < 8; i++) cv[i] = IV[i];
  size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
  for (size_t b = 0; b < nblocks; b++) {
    size_t off = b * 64, bl = (len - off < 64) ? len - off : 64;
    u32 w[16]; load_block(p + off, bl, w);
    u32 flags = 0;
    if (b == 0) flags |= XFG_B3_CHUNK_START;
    if (b == nblocks - 1) { flags |= XFG_B3_CHUNK_END; if (root) flags |= XFG_B3_ROOT; }
    u32 o[8]; compress(cv, w, chunk_counter, (u32)bl, flags, o);
    for (int i = 0; i < 8; i++) cv[i] = o[i];
  }
  for (int i = 0; i < 8; i++) out[i] = cv[i];
}
inline void parent_cv(const u32 l[8], const u32 r[8], bool root, u32 out[8]) {
  u32 w[16]; for (int i = 0; i < 8; i++) { w[i] = l[i]; w[8 + i] = r[i]; }
  compress(IV, w, 0, 64, XFG_B3_PARENT | (root ? XFG_B3_ROOT : 0), out);
}
// subtree over `len` bytes starting at chunk index `c0`; left subtree takes the largest power-of-two number of chunks
inline void subtree(const u8* p, size_t len, u64 c0, bool root, u32 out[8]) {
  if (len <= 1024) { chunk_cv(p, len, c0, root, out); return; }
  size_t chunks = (len + 1023) / 1024, left = 1; while (left * 2 < chunks) left *= 2;
  u32 l[8], r[8];
  subtree(p, left * 1024, c0, false, l);
  subtree(p + left * 1024, len - left * 1024, c0 + left, false, r);
  parent_cv(l, r, root, out);
}
}  // namespace b3

inline Digest blake3(const u8* p, size_t len) {
  u32 o[8]; b3::subtree(p, len, 0, true, o);
  Digest d; for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) d[4 * i + j] = (u8)(o[i] >> (8 * j));
  return d;
}
inline Digest blake3(const std::vector<u8>& v) { return blake3(v.data(), v.size()); }

// winter-crypto Blake3_256: hash_elements / merge / merge_with_int (A.6, D)
template <class E> inline Digest hash_elements(const E* es, size_t n) {
  std::vector<u8> b; b.reserve(n * 8 * E::DEG); for (size_t i = 0; i < n; i++) put_elem(b, es[i]); return blake3(b);
}
template <class E> inline Digest hash_elements(const std::vector<E>& es) { return hash_elements(es.data(), es.size()); }
inline Digest merge(const Digest& l, const Digest& r) { u8 b[64]; std::memcpy(b, l.data(), 32); std::memcpy(b + 32, r.data(), 32); return blake3(b, 64); }
inline Digest merge_with_int(const Digest& s, u64 v) { u8 b[40]; std::memcpy(b, s.data(), 32); for (int i = 0; i < 8; i++) b[32 + i] = (u8)(v >> (8 * i)); return blake3(b, 40); }

// ---- Keccak-256 ----
namespace kk {
static constexpr u64 RC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
  0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL,
  0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL,
  0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL,
  0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static constexpr int ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static constexpr int PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
inline u64 rotl(u64 x, int n) { return (x << n) | (x >> (64 - n)); }
inline void f1600(u64 st[25]) {
  for (int r = 0; r < 24; r++) {
    u64 bc[5];
    for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
    for (int i = 0; i < 5; i++) { u64 t = bc[(i + 4) % 5] ^ rotl(bc[(i + 1) % 5], 1); for (int j = 0; j < 25; j += 5) st[j + i] ^= t; }
    u64 t = st[1];
    for (int i = 0; i < 24; i++) { int j = PIL[i]; u64 b = st[j]; st[j] = rotl(t, ROT[i]); t = b; }
    for (int j = 0; j < 25; j += 5) { u64 b[5]; for (int i = 0; i < 5; i++) b[i] = st[j + i]; for (int i = 0; i < 5; i++) st[j + i] = b[i] ^ ((~b[(i + 1) % 5]) & b[(i + 2) % 5]); }
    st[0] ^= RC[r];
  }
}
}  // namespace kk
inline Digest keccak256(const u8* p, size_t len) {
  u64 st[25] = {0}; const size_t rate = 136;
  std::vector<u8> m(p, p + len); m.push_back(0x01); while (m.size() % rate) m.push_back(0); m.back() |= 0x80;
  for (size_t off = 0; off < m.size(); off += rate) { for (size_t i = 0; i < rate / 8; i++) st[i] ^= get_u64(&m[off + 8 * i]); kk::f1600(st); }
  Digest d; for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) d[8 * i + j] = (u8)(st[i] >> (8 * j));
  return d;
}
inline Digest keccak256(const std::vector<u8>& v) { return keccak256(v.data(), v.size()); }

}  // namespace orc
