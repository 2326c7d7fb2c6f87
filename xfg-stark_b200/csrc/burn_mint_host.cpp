// burn_mint_host.cpp — host-side mirror of the reference's application layer for the burn-mint proof.
//
// Mirrors, with the same names, argument meaning and error text:
//   XfgBurnMintProver::{prove_burn_mint (input half), validate_inputs, secret_to_field_element, compute_recipient_hash}
//       src/burn_mint_prover.rs:62-107, 132-180, 195-208, 211-221
//   XfgBurnMintAir::{compute_nullifier, compute_recipient_hash, compute_commitment, build_trace}
//       src/burn_mint_air.rs:124-133, 157-170, 174-202, 442-476
// These are three Keccak-256 calls and an O(n) fill per proof: host work in the reference and here;
// the reference re-hashes them for every evaluated row (src/burn_mint_air.rs:264, 376), this backend hoists them into the
// four constants of xfg_air_consts.
#include "burn_mint_host.hpp"
#include <cstring>

namespace xfg {

// ---- Keccak-256 (original padding 0x01 .. 0x80, rate 136), sponge over a 5x5 lane state ----
namespace {
inline uint64_t rol(uint64_t v, unsigned s) { return s ? (v << s) | (v >> (64 - s)) : v; }
void keccak_permute(uint64_t A[5][5]) {   // A[y][x]
  static const unsigned RHO[5][5] = {{0, 1, 62, 28, 27}, {36, 44, 6, 55, 20}, {3, 10, 43, 25, 39}, {41, 45, 15, 21, 8}, {18, 2, 61, 56, 14}};
  uint64_t lfsr = 1;   // round constants from the degree-8 LFSR of the Keccak specification
  for (int round = 0; round < 24; round++) {
    uint64_t C[5], B[5][5];
    for (int x = 0; x < 5; x++) C[x] = A[0][x] ^ A[1][x] ^ A[2][x] ^ A[3][x] ^ A[4][x];
    for (int x = 0; x < 5; x++) { uint64_t d = C[(x + 4) % 5] ^ rol(C[(x + 1) % 5], 1); for (int y = 0; y < 5; y++) A[y][x] ^= d; }
    for (int y = 0; y < 5; y++) for (int x = 0; x < 5; x++) B[(2 * x + 3 * y) % 5][y] = rol(A[y][x], RHO[y][x]);   // rho + pi: (x,y) -> (y, 2x+3y)
    for (int y = 0; y < 5; y++) for (int x = 0; x < 5; x++) A[y][x] = B[y][x] ^ (~B[y][(x + 1) % 5] & B[y][(x + 2) % 5]);
    uint64_t rc = 0;
    for (int j = 0; j < 7; j++) {
      if (lfsr & 1) rc ^= 1ull << ((1u << j) - 1);
      lfsr = (lfsr & 0x80) ? ((lfsr << 1) ^ 0x171) : (lfsr << 1);
    }
    A[0][0] ^= rc;
  }
}
}  // namespace

void keccak256(const uint8_t* msg, size_t len, uint8_t out[32]) {
  const size_t RATE = 136;
  uint64_t A[5][5]; std::memset(A, 0, sizeof A);
  uint8_t block[RATE];
  size_t off = 0; bool done = false;
  while (!done) {
    size_t take = len - off < RATE ? len - off : RATE;
    std::memset(block, 0, RATE); std::memcpy(block, msg + off, take); off += take;
    if (take < RATE) { block[take] ^= 0x01; block[RATE - 1] ^= 0x80; done = true; }
    for (size_t i = 0; i < RATE / 8; i++) { uint64_t w = 0; for (int b = 7; b >= 0; b--) w = (w << 8) | block[8 * i + b]; A[i / 5][i % 5] ^= w; }
    keccak_permute(A);
  }
  for (int i = 0; i < 4; i++) for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(A[0][i] >> (8 * b));
}

namespace {
struct Msg { std::string b; void le64(uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((char)(v >> (8 * i))); } void tag(const char* s) { b.append(s); }
  void raw(const uint8_t* p, size_t n) { b.append((const char*)p, n); }
  uint32_t head32(uint8_t full[32] = nullptr) const { uint8_t d[32]; keccak256((const uint8_t*)b.data(), b.size(), d); if (full) std::memcpy(full, d, 32);
    return (uint32_t)d[0] | (uint32_t)d[1] << 8 | (uint32_t)d[2] << 16 | (uint32_t)d[3] << 24; } };
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
}  // namespace

// src/burn_mint_prover.rs:211-221
uint32_t prover_compute_recipient_hash(const uint8_t* addr, size_t len) { Msg m; m.raw(addr, len); m.tag("recipient"); return m.head32(); }
// src/burn_mint_air.rs:124-133
uint64_t air_compute_nullifier(const uint64_t pi[XFG_NUM_PUB_INPUTS], uint64_t secret) { Msg m; m.le64(secret); m.tag("nullifier"); m.le64(pi[XFG_PI_BURN]); return m.head32(); }
// src/burn_mint_air.rs:174-202 (with :157-170 inlined)
uint64_t air_compute_commitment(const uint64_t pi[XFG_NUM_PUB_INPUTS], uint64_t secret) {
  Msg r; r.le64(pi[XFG_PI_RECIPIENT_HASH]); r.tag("ethereum-recipient"); r.tag("fuego-to-heat-bridge");
  uint8_t rfull[32]; r.head32(rfull);
  Msg m; m.le64(secret); m.le64(pi[XFG_PI_BURN]); m.le64(pi[XFG_PI_MINT]);
  for (int i = XFG_PI_TXP0; i <= XFG_PI_TXP3; i++) m.le64(pi[i]);
  m.raw(rfull, 32);
  m.le64(pi[XFG_PI_NETWORK_ID]); m.le64(pi[XFG_PI_TARGET_CHAIN]); m.le64(pi[XFG_PI_VERSION]);
  m.tag("heat-commitment-v1");
  return m.head32();
}

int burn_mint_pack_inputs(uint64_t burn_amount, uint64_t mint_amount, const uint8_t tx_prefix_hash[32], const uint8_t* recipient, size_t recipient_len,
                          const uint8_t* secret, size_t secret_len, uint32_t network_id, uint32_t target_chain_id, uint32_t commitment_version,
                          xfg_air_consts* out, std::string& err) {
  uint64_t legacy_txn_hash = 0; for (int b = 7; b >= 0; b--) legacy_txn_hash = (legacy_txn_hash << 8) | tx_prefix_hash[b];
  // validate_inputs (src/burn_mint_prover.rs:132-180)
  if (burn_amount != XFG_STD_BURN && burn_amount != XFG_LARGE_BURN) {
    err = "Burn amount must be exactly 0.8 XFG (8,000,000 atomic units) or 800 XFG (8,000,000,000 atomic units)"; return XFG_ERR_INVALID_INPUT; }
  if (mint_amount != burn_amount) {
    err = "Mint amount " + std::to_string(mint_amount) + " does not match burn amount " + std::to_string(burn_amount) + " for 1:1 atomic unit conversion"; return XFG_ERR_INVALID_INPUT; }
  if (legacy_txn_hash == 0) { err = "Transaction hash must be greater than 0"; return XFG_ERR_INVALID_INPUT; }
  if (recipient_len != 20) { err = "Recipient address must be exactly 20 bytes"; return XFG_ERR_INVALID_INPUT; }
  // secret_to_field_element (:195-208); the reference panics on 4..7 bytes (copies secret[..8]) - reported as an error here
  if (secret_len < 4) { err = "Secret must be at least 4 bytes"; return XFG_ERR_INVALID_INPUT; }
  if (secret_len < 8) { err = "Secret must be at least 8 bytes"; return XFG_ERR_INVALID_INPUT; }
  const uint64_t secret_element = le32(secret);
  uint64_t* pi = out->pub_inputs;
  pi[XFG_PI_BURN] = (uint32_t)burn_amount;                 // `as u32` truncations of :90-94
  pi[XFG_PI_MINT] = (uint32_t)mint_amount;
  pi[XFG_PI_TXN_HASH] = (uint32_t)legacy_txn_hash;
  pi[XFG_PI_RECIPIENT_HASH] = prover_compute_recipient_hash(recipient, recipient_len);
  pi[XFG_PI_STATE] = 0;
  for (int i = 0; i < 4; i++) pi[XFG_PI_TXP0 + i] = le32(tx_prefix_hash + 4 * i);
  pi[XFG_PI_NETWORK_ID] = network_id; pi[XFG_PI_TARGET_CHAIN] = target_chain_id; pi[XFG_PI_VERSION] = commitment_version;
  out->txn_hash = (uint32_t)pi[XFG_PI_TXN_HASH];           // src/burn_mint_air.rs:362
  out->recipient_hash = (uint32_t)pi[XFG_PI_RECIPIENT_HASH];   // :365
  out->nullifier = air_compute_nullifier(pi, secret_element);
  out->commitment = air_compute_commitment(pi, secret_element);
  return XFG_OK;
}

void burn_mint_build_trace(const xfg_air_consts* air, uint32_t n_log2, uint64_t* t) {
  const size_t n = size_t(1) << n_log2;
  const uint64_t fill[XFG_TRACE_WIDTH] = {air->pub_inputs[XFG_PI_BURN], air->pub_inputs[XFG_PI_MINT], air->pub_inputs[XFG_PI_TXN_HASH],
                                          air->pub_inputs[XFG_PI_RECIPIENT_HASH], 0, air->nullifier, air->commitment};
  for (int c = 0; c < XFG_TRACE_WIDTH; c++) for (size_t i = 0; i < n; i++) t[c * n + i] = (c == 4) ? (uint64_t)((4 * i) >> n_log2) : fill[c];
}

}  // namespace xfg
