// merkle.cuh — interface of the row-hashing / Merkle-tree kernels (see merkle.cu).
#pragma once
#include <cuda_runtime.h>
#include "state.cuh"

namespace xfg {

// Fiat-Shamir step fused into the tree kernel: the CTA that computes the root continues (warp 0) with what winter-prover does with a commitment
// (`commit_trace` / `commit_constraints` / `commit_fri_layer`: reseed the coin with the root, then the draws that follow it), instead of a
// dependent single-warp launch per tree.
struct RootStep {
  int kind;                 // 0 none; 1 trace root: coin seed (A.4), reseed, `count` coefficients -> out; 2 constraint root: z, z g; 3 FRI layer `layer`: alpha
  int D; ProofState* ps; u64 (*out)[2]; u32 count; u64 g_n; u32 layer;
  unsigned* ticket;         // zero-initialised "last CTA" counter for the fused upper-tree kernel (null: heap slot 0 of the tree, cleared by a memset)
};

// Heap layout of a tree over M leaves: tree[M + i] = leaf i, tree[i] = BLAKE3(tree[2i] || tree[2i+1]) for 1 <= i < M,
// tree[1] = root (winter-crypto MerkleTree::new keeps the same `nodes` numbering, A.7).

// Leaves of a coset-major matrix (NL limb arrays of 8 cosets x n points each; LDE row i = 8m + k lives at [k*n + m]):
// hashes the 8 rows of each m and the 3 tree levels above them.  tree has 2*8n digests.
// leaf hashing + 3 levels only (levels N .. N/8); merkle_commit_rows = this + merkle_build_upper(tree, n)
void launch_commit_rows(cudaStream_t st, const u64* data, size_t limb_stride, int num_limbs, u32 ln, Digest* tree);
// rows of up to 128 limbs (one BLAKE3 chunk), streamed block by block; same outputs as launch_commit_rows
void launch_commit_rows_wide(cudaStream_t st, const u64* data, size_t limb_stride, u32 num_limbs, u32 ln, Digest* tree);
void merkle_commit_rows(cudaStream_t st, const u64* data, size_t limb_stride, int num_limbs, u32 ln, Digest* tree);
// Completes the tree above a fully written level of M nodes (heap range [M, 2M)).
void merkle_build_upper(cudaStream_t st, Digest* tree, size_t M, const RootStep* step = nullptr);
// hash_elements of `count` row-major rows of `limbs` (1, 2, 7, 8, 16) elements
void launch_hash_rows(cudaStream_t st, const u64* rows, size_t count, int limbs, Digest* out);

}  // namespace xfg
