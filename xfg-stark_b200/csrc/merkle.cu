// merkle.cu — BLAKE3 row hashing fused with the first tree levels, and the upper Merkle levels, for sm_100a.
//
// Replaces winter-prover 0.8.3 `RowMatrix::commit_to_rows` + winter-crypto `MerkleTree::new`
// (reached from DefaultTraceLde::new at src/burn_mint_air.rs:513 and from
// `Prover::build_constraint_commitment`).  These kernels are integer-pipe bound (one BLAKE3 compression is ~700 32-bit
// ALU ops against 56-64 input bytes), so every lane keeps a full compression busy: a thread hashes the 8 LDE rows of one
// trace step (8 leaves = one complete 3-level subtree in the coset-major layout) and reduces them itself; the levels above
// are reduced 8 -> 1 per thread as well, never by a shrinking tree inside a warp.
#include "merkle.cuh"
#include "b3_rolled.cuh"
#include "launch.cuh"

namespace xfg {

#ifndef XFG_COMMIT_ROLLED
#define XFG_COMMIT_ROLLED 1
#endif
#if XFG_COMMIT_ROLLED
// One inlined leaf compression and one inlined node compression, each in a rolled loop: the unrolled kernel (15 compressions, 144 KB of
// SASS; 8 per iteration after the first rolling, 88 KB) stalled on instruction fetch (ncu: no_instruction was its top stall).  The 8 leaf
// digests of the thread are staged in shared memory ([slot][word][thread], conflict-free) and reduced in place: node j of a level reads
// slots 2j, 2j+1 and writes slot j, which the in-order loop has already consumed.
template <int NL>
__global__ void __launch_bounds__(128) commit_rows_kernel(const u64* __restrict__ data, size_t limb_stride, u32 ln, Digest* __restrict__ tree) {
  __shared__ u32 sd[8][8][128];
  const size_t n = size_t(1) << ln, N = n * 8;
  const u32 tid = threadIdx.x;
  const size_t m = (size_t)blockIdx.x * blockDim.x + tid;
  if (m >= n) return;
#pragma unroll 1
  for (int k = 0; k < 8; k++) {
    u64 limbs[NL];
#pragma unroll
    for (int j = 0; j < NL; j++) limbs[j] = data[j * limb_stride + (size_t)k * n + m];
    const Digest d = b3_hash_limbs<NL>(limbs);
    store_digest(tree + N + 8 * m + k, d);
#pragma unroll
    for (int i = 0; i < 8; i++) sd[k][i][tid] = d.w[i];
  }
#pragma unroll 1
  for (int it = 0; it < 7; it++) {
    const int lvl = it < 4 ? 0 : it < 6 ? 1 : 2, j = it - (lvl == 0 ? 0 : lvl == 1 ? 4 : 6), cnt = 4 >> lvl;
    Digest l, r;
#pragma unroll
    for (int i = 0; i < 8; i++) { l.w[i] = sd[2 * j][i][tid]; r.w[i] = sd[2 * j + 1][i][tid]; }
    const Digest d = b3_merge(l, r);
    store_digest(tree + (N >> (lvl + 1)) + (size_t)cnt * m + j, d);
#pragma unroll
    for (int i = 0; i < 8; i++) sd[j][i][tid] = d.w[i];
  }
}
#else
template <int NL>
__global__ void __launch_bounds__(128) commit_rows_kernel(const u64* __restrict__ data, size_t limb_stride, u32 ln, Digest* __restrict__ tree) {
  const size_t n = size_t(1) << ln, N = n * 8;
  const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  // The two halves (cosets 0-3, 4-7) run as a rolled loop: the fully unrolled body is 15 compressions = 144 KB of SASS and stalled on
  // instruction fetch (ncu: no_instruction 8.8 warps per issue); 8 compressions per iteration fit the instruction cache.
  Digest top[2];
#pragma unroll 1
  for (int half = 0; half < 2; half++) {
    Digest l1[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int kp = 2 * half + q;
      Digest d[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int k = 2 * kp + h;
        u64 limbs[NL];
#pragma unroll
        for (int j = 0; j < NL; j++) limbs[j] = data[j * limb_stride + (size_t)k * n + m];
        d[h] = b3_hash_limbs<NL>(limbs);
        store_digest(tree + N + 8 * m + k, d[h]);
      }
      l1[q] = b3_merge(d[0], d[1]);
      store_digest(tree + N / 2 + 4 * m + kp, l1[q]);
    }
    const Digest r = b3_merge(l1[0], l1[1]);
    store_digest(tree + N / 4 + 2 * m + half, r);
    if (half == 0) top[0] = r; else top[1] = r;
  }
  store_digest(tree + N / 8 + m, b3_merge(top[0], top[1]));
}

#endif

// Wide rows (config 5: W = 64 columns): same thread shape, but the row is streamed through the BLAKE3 chunk 8 limbs (one
// 64-byte block) at a time instead of being held in registers.  data[j*limb_stride + k*n + m], num_limbs <= 128.
__global__ void __launch_bounds__(128) commit_rows_wide_kernel(const u64* __restrict__ data, size_t limb_stride, u32 num_limbs, u32 ln, Digest* __restrict__ tree) {
  const size_t n = size_t(1) << ln, N = n * 8;
  const size_t m = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const u32 nb = (num_limbs + 7) / 8;
  Digest l1[4];
#pragma unroll 1
  for (int kp = 0; kp < 4; kp++) {
    Digest d[2];
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      const int k = 2 * kp + h;
      const u64* row = data + (size_t)k * n + m;
      u32 cv[8]; b3_iv(cv);
      for (u32 b = 0; b < nb; b++) {
        u32 msg[16];
#pragma unroll
        for (int i = 0; i < 8; i++) { const u32 li = b * 8 + i; const u64 v = li < num_limbs ? row[(size_t)li * limb_stride] : 0; msg[2 * i] = (u32)v; msg[2 * i + 1] = (u32)(v >> 32); }
        const u32 rem = num_limbs - b * 8, len = rem >= 8 ? 64 : rem * 8;
        const u32 flags = (b == 0 ? XFG_B3_CHUNK_START : 0) | (b == nb - 1 ? (XFG_B3_CHUNK_END | XFG_B3_ROOT) : 0);
        b3_compress(cv, msg, len, flags, b == nb - 1 ? d[h].w : cv);
      }
      store_digest(tree + N + 8 * m + k, d[h]);
    }
    l1[kp] = b3_merge(d[0], d[1]);
    store_digest(tree + N / 2 + 4 * m + kp, l1[kp]);
  }
  Digest a = b3_merge(l1[0], l1[1]), b = b3_merge(l1[2], l1[3]);
  store_digest(tree + N / 4 + 2 * m, a); store_digest(tree + N / 4 + 2 * m + 1, b);
  store_digest(tree + N / 8 + m, b3_merge(a, b));
}
void launch_commit_rows_wide(cudaStream_t st, const u64* data, size_t limb_stride, u32 num_limbs, u32 ln, Digest* tree) {
  const size_t n = size_t(1) << ln;
  commit_rows_wide_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(data, limb_stride, num_limbs, ln, tree);
  XFG_LAUNCHED(1);
}

// level of M nodes at [M, 2M) -> levels M/2, M/4, M/8; one thread per 8 children.  One inlined compression in a rolled loop of 7 (see
// commit_rows_kernel): the first 4 nodes read their children from the heap, the others from the shared-memory staging.
#if XFG_COMMIT_ROLLED
__global__ void __launch_bounds__(128) tree_reduce8_kernel(Digest* __restrict__ tree, size_t M) {
  __shared__ u32 sd[4][8][128];
  const u32 tid = threadIdx.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + tid;
  if (t >= M / 8) return;
#pragma unroll 1
  for (int it = 0; it < 7; it++) {
    const int lvl = it < 4 ? 0 : it < 6 ? 1 : 2, j = it - (lvl == 0 ? 0 : lvl == 1 ? 4 : 6), cnt = 4 >> lvl;
    Digest l, r;
    if (lvl == 0) { l = load_digest(tree + M + 8 * t + 2 * j); r = load_digest(tree + M + 8 * t + 2 * j + 1); }
    else {
#pragma unroll
      for (int i = 0; i < 8; i++) { l.w[i] = sd[2 * j][i][tid]; r.w[i] = sd[2 * j + 1][i][tid]; }
    }
    const Digest d = b3_merge(l, r);
    store_digest(tree + (M >> (lvl + 1)) + (size_t)cnt * t + j, d);
#pragma unroll
    for (int i = 0; i < 8; i++) sd[j][i][tid] = d.w[i];
  }
}
#else
__global__ void __launch_bounds__(128) tree_reduce8_kernel(Digest* __restrict__ tree, size_t M) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M / 8) return;
  Digest l1[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    Digest x = load_digest(tree + M + 8 * t + 2 * j), y = load_digest(tree + M + 8 * t + 2 * j + 1);
    l1[j] = b3_merge(x, y);
    store_digest(tree + M / 2 + 4 * t + j, l1[j]);
  }
  Digest a = b3_merge(l1[0], l1[1]), b = b3_merge(l1[2], l1[3]);
  store_digest(tree + M / 4 + 2 * t, a); store_digest(tree + M / 4 + 2 * t + 1, b);
  store_digest(tree + M / 8 + t, b3_merge(a, b));
}

#endif

// warp 0 of the CTA that has just written the root (tree[1]): the transcript step that consumes it (see RootStep)
static __device__ void root_step(const RootStep& rs, const Digest* tree) {
  if (rs.kind == 0 || threadIdx.x >= 32) return;
  __syncwarp();
  const Digest root = load_digest(tree + 1);
  ProofState* ps = rs.ps;
  Coin c;
  if (rs.kind == 1) { c.seed = r_hash_limbs(ps->seed_limbs, (int)ps->seed_count); c.counter = 0; }      // coin = hash_elements(context || public inputs)
  else c = coin_load(ps);
  r_reseed(c, root);
  bool ok;
  if (rs.kind == 1) ok = rs.D == 2 ? r_draw_many<2>(c, rs.count, rs.out) : r_draw_many<1>(c, rs.count, rs.out);
  else if (rs.kind == 2) ok = rs.D == 2 ? r_draw_many<2>(c, 1, &ps->z) : r_draw_many<1>(c, 1, &ps->z);
  else ok = rs.D == 2 ? r_draw_many<2>(c, 1, &ps->alphas[rs.layer]) : r_draw_many<1>(c, 1, &ps->alphas[rs.layer]);
  if (lane_id() == 0) {
    if (rs.kind == 1) ps->trace_root = root;
    else if (rs.kind == 2) {
      ps->constraint_root = root;
      if (rs.D == 2) stx<2>(ps->zg, mul_base(ldx<2>(ps->z), rs.g_n)); else stx<1>(ps->zg, mul_base(ldx<1>(ps->z), rs.g_n));
    } else ps->fri_roots[rs.layer] = root;
    if (!ok) ps->error_flags |= ERR_FLAG_COIN;
  }
  coin_store(ps, c);
}

// one CTA finishes the tree from a level of M <= 2048 nodes up to the root.  Every level is written to the heap (authentication
// paths need all nodes) but the next level reads its children from shared memory (ping-pong buffers), so a level costs one
// compression latency + a barrier instead of a global-memory round trip; the last levels run inside one warp.
__global__ void __launch_bounds__(1024) tree_top_kernel(Digest* __restrict__ tree, u32 M, RootStep rs) {
  __shared__ Digest bufA[1024], bufB[512];
  Digest* out = bufA; Digest* in = nullptr;
  for (u32 lvl = M / 2; lvl >= 1; lvl >>= 1) {
    for (u32 i = threadIdx.x; i < lvl; i += blockDim.x) {
      const Digest x = in ? in[2 * i] : load_digest(tree + 2 * (lvl + i)), y = in ? in[2 * i + 1] : load_digest(tree + 2 * (lvl + i) + 1);
      const Digest d = b3_merge(x, y);
      out[i] = d; store_digest(tree + lvl + i, d);
    }
    if (lvl > 16) __syncthreads(); else __syncwarp();     // levels of <= 16 nodes are produced and consumed by warp 0 only
    in = out; out = (out == bufA) ? bufB : bufA;
  }
  root_step(rs, tree);
}

// hash_elements of row-major rows (stage entry point xfg_hash_rows; the pipeline hashes rows inside its fused kernels)
template <int NL>
__global__ void __launch_bounds__(128) hash_rows_kernel(const u64* __restrict__ rows, size_t count, Digest* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  u64 limbs[NL];
#pragma unroll
  for (int j = 0; j < NL; j++) limbs[j] = rows[i * NL + j];
  store_digest(out + i, b3_hash_limbs<NL>(limbs));
}
void launch_hash_rows(cudaStream_t st, const u64* rows, size_t count, int limbs, Digest* out) {
  const unsigned blocks = (unsigned)((count + 127) / 128);
  switch (limbs) {
    case 1: hash_rows_kernel<1><<<blocks, 128, 0, st>>>(rows, count, out); break;
    case 2: hash_rows_kernel<2><<<blocks, 128, 0, st>>>(rows, count, out); break;
    case 7: hash_rows_kernel<7><<<blocks, 128, 0, st>>>(rows, count, out); break;
    case 8: hash_rows_kernel<8><<<blocks, 128, 0, st>>>(rows, count, out); break;
    case 16: hash_rows_kernel<16><<<blocks, 128, 0, st>>>(rows, count, out); break;
    default: return;
  }
  XFG_LAUNCHED(1);
}

void launch_commit_rows(cudaStream_t st, const u64* data, size_t limb_stride, int num_limbs, u32 ln, Digest* tree) {
  const size_t n = size_t(1) << ln; const unsigned blocks = (unsigned)((n + 127) / 128);
  switch (num_limbs) {
    case 1: commit_rows_kernel<1><<<blocks, 128, 0, st>>>(data, limb_stride, ln, tree); break;
    case 2: commit_rows_kernel<2><<<blocks, 128, 0, st>>>(data, limb_stride, ln, tree); break;
    case 7: commit_rows_kernel<7><<<blocks, 128, 0, st>>>(data, limb_stride, ln, tree); break;
    default: return;   // callers only pass 1, 2 or XFG_TRACE_WIDTH
  }
  XFG_LAUNCHED(1);
}
void merkle_commit_rows(cudaStream_t st, const u64* data, size_t limb_stride, int num_limbs, u32 ln, Digest* tree) {
  launch_commit_rows(st, data, limb_stride, num_limbs, ln, tree);
  merkle_build_upper(st, tree, size_t(1) << ln);
}

// Fused upper tree for 2048 < M <= 2^20: CTA b reduces the 1024 nodes [M + 1024 b, M + 1024 (b+1)) to the node M/1024 + b through
// shared memory (10 levels, every level also written to the heap); the last CTA to finish (ticket counter kept in the unused
// heap slot 0, zeroed by the caller) then reduces the M/1024 subtree roots to the root.  One launch instead of 3-4.
__device__ __forceinline__ void smem_tree_levels(Digest* __restrict__ tree, size_t first_parent_level, size_t offset, u32 count, Digest* bufA, Digest* bufB) {
  // reduces `count` (power of two, <= 2048) nodes at heap indices [2*(first_parent_level + offset) ...) ; parents of level size L live at L + offset_L
  Digest* out = bufA; Digest* in = nullptr;
  size_t lvl = first_parent_level, off = offset;
  for (u32 c = count / 2; c >= 1; c >>= 1, lvl >>= 1, off >>= 1) {
    for (u32 i = threadIdx.x; i < c; i += blockDim.x) {
      const size_t node = lvl + off + i;
      const Digest x = in ? in[2 * i] : load_digest(tree + 2 * node), y = in ? in[2 * i + 1] : load_digest(tree + 2 * node + 1);
      const Digest d = b3_merge(x, y);
      out[i] = d; store_digest(tree + node, d);
    }
    if (c > 16) __syncthreads(); else __syncwarp();
    in = out; out = (out == bufA) ? bufB : bufA;
  }
}
__global__ void __launch_bounds__(512) tree_upper_fused_kernel(Digest* __restrict__ tree, size_t M, RootStep rs) {
  __shared__ Digest bufA[512], bufB[256];
  __shared__ bool is_last;
  smem_tree_levels(tree, M / 2, (size_t)blockIdx.x * 512, 1024, bufA, bufB);     // -> node M/1024 + blockIdx.x
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ticket = rs.ticket ? rs.ticket : reinterpret_cast<unsigned*>(tree);
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (is_last) *ticket = 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const u32 roots = (u32)(M / 1024);                                              // <= 1024
  smem_tree_levels(tree, roots / 2, 0, roots, bufA, bufB);
  root_step(rs, tree);
}

void merkle_build_upper(cudaStream_t st, Digest* tree, size_t M, const RootStep* step) {
  RootStep rs{}; if (step) rs = *step;
  // big levels: 8 -> 1 per thread (every lane busy for 7 compressions); from 2^17 nodes down: one fused launch
  while (M > (size_t(1) << 17)) {
    const size_t t = M / 8;
    tree_reduce8_kernel<<<(unsigned)((t + 127) / 128), 128, 0, st>>>(tree, M);
    XFG_LAUNCHED(1);
    M /= 8;
  }
  if (M > 2048) {
    if (!rs.ticket) cudaMemsetAsync(tree, 0, sizeof(unsigned), st);
    tree_upper_fused_kernel<<<(unsigned)(M / 1024), 512, 0, st>>>(tree, M, rs); XFG_LAUNCHED(1);
    return;
  }
  if (M >= 2) { tree_top_kernel<<<1, 1024, 0, st>>>(tree, (u32)M, rs); XFG_LAUNCHED(1); }
}

}  // namespace xfg
