// transcript.cuh — launchers of the device-side Fiat-Shamir channel and query gather (see transcript.cu).
#pragma once
#include <cuda_runtime.h>
#include "state.cuh"

namespace xfg {

// One opened commitment: rows of J x limbs elements at the queried positions plus `depth` sibling digests per position.
// Element (j, l) of position p is src[l*limb_stride + addr(p + j*R)], addr = coset-major when coset_n != 0.
struct GatherTask {
  const u64* src; const Digest* tree;
  u64 limb_stride, coset_n, R, M;      // M = number of leaves (heap offset of the leaf level)
  u32 J, limbs, depth; int fri_layer;  // fri_layer < 0: the LDE query positions, else the folded positions of that layer
  u64 rows_off, paths_off;             // offsets into the material buffer, in u64 units
};
struct GatherTasks { GatherTask t[2 + MAX_LAYERS]; u32 count; };

void launch_ood_finish(cudaStream_t st, int D, ProofState* ps, const u64* partial, u32 nb);
void launch_grind(cudaStream_t st, ProofState* ps, u32 grinding);
void launch_positions(cudaStream_t st, ProofState* ps, u32 num_queries, u32 lN, u32 num_layers);
void launch_gather(cudaStream_t st, const GatherTasks& tasks, const ProofState* ps, u64* out);

}  // namespace xfg
