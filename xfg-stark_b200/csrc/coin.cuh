// coin.cuh — warp-parallel Fiat-Shamir coin shared by the prover transcript (transcript.cu) and the batch verifier (verify.cu).
//
// Replaces winter-crypto 0.8.3 `DefaultRandomCoin::{reseed, draw}` on the device.
#pragma once
#include "state.cuh"

namespace xfg {

// ---- coin (one warp per transcript kernel; lanes hash candidate counters in parallel, results are identical to the serial
// winter-crypto loop because the k-th accepted value is the k-th valid candidate in counter order) ----
struct Coin { Digest seed; u64 counter; };
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ Digest bcast_digest(const Digest& d, int src) { Digest r; for (int i = 0; i < 8; i++) r.w[i] = __shfl_sync(0xFFFFFFFFu, d.w[i], src); return r; }
__device__ __forceinline__ Coin coin_load(const ProofState* ps) { Coin c; c.seed = ps->seed; c.counter = ps->counter; return c; }
__device__ __forceinline__ void coin_store(ProofState* ps, const Coin& c) { if (lane_id() == 0) { ps->seed = c.seed; ps->counter = c.counter; } }
// reseed(d): seed = BLAKE3(seed || d), counter = 0 (every lane computes the same value; the state stays warp-uniform)
__device__ __forceinline__ void coin_reseed(Coin& c, const Digest& d) { c.seed = b3_merge(c.seed, d); c.counter = 0; }
// `count` consecutive draw::<E>() calls: first 8*D bytes of next(); every limb must be canonical, else the candidate is skipped (A.5)
template <int D> __device__ bool coin_draw_many(Coin& c, u32 count, u64 (*out)[2]) {
  u32 got = 0;
  for (int round = 0; round < 40 && got < count; round++) {
    const Digest d = b3_merge_int(c.seed, c.counter + 1 + lane_id());
    const u64 v0 = (u64)d.w[0] | ((u64)d.w[1] << 32), v1 = (u64)d.w[2] | ((u64)d.w[3] << 32);
    const bool valid = v0 < GL_P && (D == 1 || v1 < GL_P);
    const u32 mask = __ballot_sync(0xFFFFFFFFu, valid), rank = __popc(mask & ((1u << lane_id()) - 1)), need = count - got;
    if (valid && rank < need) { out[got + rank][0] = v0; out[got + rank][1] = D == 2 ? v1 : 0; }
    const u32 nvalid = __popc(mask);
    if (nvalid >= need) { c.counter += __fns(mask, 0, need) + 1; got = count; }     // counter stops at the last consumed candidate
    else { c.counter += 32; got += nvalid; }
  }
  __syncwarp();
  return got == count;
}
template <int D> __device__ __forceinline__ Ext<D> ldx(const u64* p) { return Ext<D>(p[0], p[1]); }
template <int D> __device__ __forceinline__ void stx(u64* p, Ext<D> v) { p[0] = v.limb(0); p[1] = D == 2 ? v.limb(1) : 0; }

}  // namespace xfg
