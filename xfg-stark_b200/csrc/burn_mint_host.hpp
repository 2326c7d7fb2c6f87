// burn_mint_host.hpp — host-side mirror of XfgBurnMintProver / XfgBurnMintAir input handling (see burn_mint_host.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include "../../include/xfg_stark.h"

namespace xfg {

void keccak256(const uint8_t* msg, size_t len, uint8_t out[32]);
uint32_t prover_compute_recipient_hash(const uint8_t* addr, size_t len);
uint64_t air_compute_nullifier(const uint64_t pi[XFG_NUM_PUB_INPUTS], uint64_t secret);
uint64_t air_compute_commitment(const uint64_t pi[XFG_NUM_PUB_INPUTS], uint64_t secret);
int burn_mint_pack_inputs(uint64_t burn_amount, uint64_t mint_amount, const uint8_t tx_prefix_hash[32], const uint8_t* recipient, size_t recipient_len,
                          const uint8_t* secret, size_t secret_len, uint32_t network_id, uint32_t target_chain_id, uint32_t commitment_version,
                          xfg_air_consts* out, std::string& err);
void burn_mint_build_trace(const xfg_air_consts* air, uint32_t n_log2, uint64_t* trace_colmajor);

}  // namespace xfg
