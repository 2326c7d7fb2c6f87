// verify.cuh — batch verification of burn-mint proofs on the device (see verify.cu).
#pragma once
#include <cuda_runtime.h>
#include "state.cuh"
#include "stark_kernels.cuh"

namespace xfg {

// One commitment opening of a proof: the queried values and the BatchMerkleProof node vectors, located inside the raw proof bytes
// (offsets in bytes from the start of the proof; nothing is aligned).  `idx_off` points into the batch's vector index:
// one (byte offset of the first digest, number of digests) pair per node vector, written by the host's structural parse.
struct VerifyOpening { u32 vals_off, vals_count /* u64 limbs */, idx_off, num_vecs; };

// One proof, parsed (structure only) by the host; every cryptographic and algebraic check happens in verify_kernel.
struct VerifyRec {
  u32 ln, D, num_layers, num_unique, num_queries, grinding, rem_len, host_status;   // host_status != 0: rejected while parsing
  u64 base;                                        // byte offset of the proof in the batch data buffer (8-byte aligned)
  u64 nonce;
  u64 seed_limbs[8 + XFG_NUM_PUB_INPUTS];          // Context::to_elements() || public inputs (A.4)
  AirParams air;
  Digest commitments[3 + MAX_LAYERS];              // trace root, constraint root, FRI layer roots, remainder commitment
  u64 ood_frame[2 * XFG_TRACE_WIDTH][2]; u64 hz[2];
  u64 remainder[MAX_REMAINDER][2];
  VerifyOpening op[2 + MAX_LAYERS];                // trace, constraint, FRI layers
  FriConsts fc;
};

void launch_verify(cudaStream_t st, const VerifyRec* recs, const u8* data, const uint2* vec_index, int* results, u32 count);

}  // namespace xfg
