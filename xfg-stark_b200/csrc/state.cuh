// state.cuh — device-resident proof state: the Fiat-Shamir transcript and every value derived from it.
//
// Plays the role of winter-prover 0.8.3 `ProverChannel` + `DefaultRandomCoin`: the coin and
// all drawn challenges stay in HBM so that the whole proof is one dependent chain of kernel launches with no host
// round trip until the final copy-out.
#pragma once
#include "blake3.cuh"

namespace xfg {

static constexpr int MAX_Q = XFG_MAX_QUERIES;          // 255
static constexpr int MAX_LAYERS = XFG_MAX_FRI_LAYERS;  // 16
static constexpr int MAX_REMAINDER = 256;              // (fri_remainder_max_degree + 1) <= 256 coefficients
#ifndef XFG_OOD_BLOCKS
#define XFG_OOD_BLOCKS 64      // blocks per polynomial of the out-of-domain evaluation (128 was measured: 0.097 -> 0.111 ms at 2^20)
#endif
static constexpr int OOD_MAX_BLOCKS = XFG_OOD_BLOCKS;
static constexpr int OOD_MAX_STEPS = 1024;             // n / (OOD_MAX_BLOCKS * 256) at n = 2^24
static constexpr int NUM_OOD_POLYS = XFG_TRACE_WIDTH + 2;   // 7 trace polys + up to 2 limb polys of H

enum : u32 { ERR_FLAG_DEGREE = 1u, ERR_FLAG_COIN = 2u, ERR_FLAG_NONCANONICAL = 4u };

static constexpr int MAX_SEED_LIMBS = 8 + XFG_AIR_MAX_PUB_INPUTS;   // Context::to_elements (8) + public inputs

struct ProofState {
  // ---- per-proof inputs: this block is written by ONE host->device copy from the slot's pinned mirror at the start of every proof (no
  // seeding kernel): the coin seed elements (A.4), and the reset values of the error flags and of the grinding nonce
  u64 seed_limbs[MAX_SEED_LIMBS]; u32 seed_count; u32 error_flags; unsigned long long nonce;   // nonce: ~0 until the grinding search has found it
  u32 tickets[16];          // "last CTA finishes the tree" counters of the upper-tree kernels (one per tree of the proof), zeroed by the same copy
  // coin
  Digest seed; u64 counter;
  // commitments
  Digest trace_root, constraint_root, fri_roots[MAX_LAYERS], remainder_commitment;
  // challenges (limbs [2] even when the extension degree is 1)
  u64 tcoef[XFG_NUM_TRANSITION][2], bcoef[XFG_NUM_ASSERTIONS][2];
  u64 z[2], zg[2];
  u64 ood_frame[2 * XFG_TRACE_WIDTH][2];   // T_0(z), T_0(zg), T_1(z), ... (A.9 interleaving)
  u64 hz[2];
  u64 dcoef[XFG_TRACE_WIDTH + 1][2];
  u64 deep_c1[2], deep_c2[2];              // sum_j gamma_j T_j(z) + delta H(z)   and   sum_j gamma_j T_j(zg)
  u64 alphas[MAX_LAYERS][2];
  u64 remainder[MAX_REMAINDER][2]; u32 remainder_len;
  // queries (the grinding nonce lives in the init block above)
  u32 num_positions; u32 positions[MAX_Q];
  u32 fri_num_positions[MAX_LAYERS]; u32 fri_positions[MAX_LAYERS][MAX_Q];
#ifdef XFG_TAIL_CLOCKS
  long long dbg_clk[48];   // phase timestamps of the fused tail kernel (debug builds only)
#endif
};
static constexpr size_t PROOF_INIT_BYTES = sizeof(u64) * MAX_SEED_LIMBS + 16 + 64;
static_assert(PROOF_INIT_BYTES == offsetof(ProofState, seed), "init block layout");

// AIR constants and boundary values (src/burn_mint_air.rs:335-395), passed by value to the constraint kernel
struct AirParams {
  u64 txn, rcpt, nullifier, commitment;     // transition-constraint constants
  u64 assert0[XFG_TRACE_WIDTH];             // step-0 assertion values, columns 0..6
  u64 g_last;                               // g^(n-1): transition exemption point and last-step assertion point
};

}  // namespace xfg
