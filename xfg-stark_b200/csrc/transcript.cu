// transcript.cu — the Fiat-Shamir channel on the device: coin seeding / reseeding / draws, grinding, query positions and
// the gather of queried rows and authentication nodes.
//
// Replaces winter-prover 0.8.3 `ProverChannel::{new, commit_trace, commit_constraints, get_*_coeffs, get_ood_point,
// send_ood_*, commit_fri_layer, draw_fri_alpha, grind_query_seed, get_query_positions}`, winter-crypto
// `DefaultRandomCoin`, winter-fri `fold_positions` and the row/path collection of `DefaultTraceLde::query`,
// `ConstraintCommitment::query`, `FriProver::build_proof` (A.4, A.5, A.9-A.11).
// The serial steps are single-warp kernels (lanes hash candidate counters in parallel) so the proof needs no host round trip.
#include "transcript.cuh"
#include "b3_rolled.cuh"
#include "launch.cuh"

namespace xfg {

// The root-consuming steps (commit_trace / commit_constraints / commit_fri_layer and the draws that follow them) run inside the tree kernels
// (merkle.cu: RootStep); the remainder commitment, grinding and positions of ordinary proofs inside fri_tail.cu.
// sums the OOD partials, sends the frame and H(z) to the coin, draws the DEEP coefficients (A.9)
// Block of (7 + D) warps: warp p first sums the partials of polynomial p (all loads of a lane issued together: one memory latency
// instead of one per sum), then warp 0 alone continues with the transcript.
template <int D> __global__ void __launch_bounds__(32 * NUM_OOD_POLYS) ood_finish_kernel(ProofState* ps, const u64* __restrict__ partial, u32 nb) {
  __shared__ u64 sums[NUM_OOD_POLYS][2][2];
  __shared__ u64 cpart[2][XFG_TRACE_WIDTH + 1][2];
  {
    const u32 p = threadIdx.x >> 5;        // exact arithmetic: any summation order gives the same field element
    u64 v[OOD_MAX_BLOCKS / 32][4];
#pragma unroll
    for (u32 i = 0; i < OOD_MAX_BLOCKS / 32; i++) { const u32 b = lane_id() + 32 * i;
#pragma unroll
      for (u32 q = 0; q < 4; q++) v[i][q] = b < nb ? partial[((size_t)p * nb + b) * 4 + q] : 0; }
#pragma unroll
    for (u32 q = 0; q < 4; q++) {
      u64 s = 0;
#pragma unroll
      for (u32 i = 0; i < OOD_MAX_BLOCKS / 32; i++) s = gl_add(s, v[i][q]);
      for (int o = 16; o > 0; o >>= 1) s = gl_add(s, __shfl_xor_sync(0xFFFFFFFFu, s, o));
      if (lane_id() == 0) sums[p][q >> 1][q & 1] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  Coin c = coin_load(ps);
  u64 limbs[2 * XFG_TRACE_WIDTH * D];
#pragma unroll
  for (int j = 0; j < XFG_TRACE_WIDTH; j++)
#pragma unroll
    for (int w = 0; w < 2; w++)
#pragma unroll
      for (int l = 0; l < D; l++) limbs[(2 * j + w) * D + l] = sums[j][w][l];   // interleaved per column (A.9)
  r_reseed(c, r_hash_limbs(limbs, 2 * XFG_TRACE_WIDTH * D));
  // H(z) = P_limb0(z) + phi * P_limb1(z), phi = (0,1): (a0,a1) * phi = (-2 a1, a0 + a1)
  Ext<D> hz = ldx<D>(sums[XFG_TRACE_WIDTH][0]);
  if (D == 2) { const u64 a0 = sums[XFG_TRACE_WIDTH + 1][0][0], a1 = sums[XFG_TRACE_WIDTH + 1][0][1]; hz = hz + Ext<D>(gl_neg(gl_dbl(a1)), gl_add(a0, a1)); }
  u64 hl[2] = {hz.limb(0), hz.limb(1)};
  r_reseed(c, r_hash_limbs(hl, D));
  const bool ok = r_draw_many<D>(c, XFG_TRACE_WIDTH + 1, ps->dcoef);      // 7 trace coefficients, then 1 composition column
  __syncwarp();
  // C1 = sum_j gamma_j T_j(z) + delta H(z), C2 = sum_j gamma_j T_j(zg): lane (w, j) computes one product, lane 0 adds them up
  if (lane_id() < 2 * (XFG_TRACE_WIDTH + 1)) {
    const u32 w = lane_id() / (XFG_TRACE_WIDTH + 1), j = lane_id() % (XFG_TRACE_WIDTH + 1);
    Ext<D> t;
    if (j < XFG_TRACE_WIDTH) t = ldx<D>(ps->dcoef[j]) * ldx<D>(sums[j][w]);
    else if (w == 0) t = ldx<D>(ps->dcoef[XFG_TRACE_WIDTH]) * hz;
    stx<D>(cpart[w][j], t);
  }
  __syncwarp();
  if (lane_id() == 0) {
    for (int j = 0; j < XFG_TRACE_WIDTH; j++) for (int w = 0; w < 2; w++) { ps->ood_frame[2 * j + w][0] = sums[j][w][0]; ps->ood_frame[2 * j + w][1] = D == 2 ? sums[j][w][1] : 0; }
    stx<D>(ps->hz, hz);
    if (!ok) ps->error_flags |= ERR_FLAG_COIN;
    Ext<D> c1, c2;
    for (int j = 0; j <= XFG_TRACE_WIDTH; j++) { c1 = c1 + ldx<D>(cpart[0][j]); c2 = c2 + ldx<D>(cpart[1][j]); }
    stx<D>(ps->deep_c1, c1); stx<D>(ps->deep_c2, c2);
  }
  coin_store(ps, c);
}
// grinding: smallest nonce >= 1 with trailing_zeros(LE head of BLAKE3(seed || nonce)) >= grinding_factor (A.5).
// Thread g tests g+1, g+1+TOT, ...; it stops as soon as its next candidate exceeds the best found so far, so every
// smaller candidate is always tested and the result is the serial minimum.
__global__ void __launch_bounds__(256) grind_kernel(ProofState* ps, u32 grinding) {
  const Digest seed = ps->seed;
  const u64 TOT = (u64)gridDim.x * blockDim.x, g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 mask = grinding >= 64 ? ~0ull : ((1ull << grinding) - 1);
  for (u64 nonce = g + 1;; nonce += TOT) {
    if (nonce > *(volatile unsigned long long*)&ps->nonce) break;
    Digest d = b3_merge_int(seed, nonce);
    u64 head = (u64)d.w[0] | ((u64)d.w[1] << 32);
    if ((head & mask) == 0) { atomicMin(&ps->nonce, (unsigned long long)nonce); break; }
  }
}
// draw_integers(q, N, nonce) -> sort -> dedup; then fold_positions per FRI layer (A.5, A.10).  One warp: lanes hash the q
// counters in parallel, rank-sort in shared memory, and compact with ballots (order-preserving, as the reference).
__global__ void __launch_bounds__(32) positions_kernel(ProofState* ps, u32 num_queries, u32 lN, u32 num_layers) {
  __shared__ u32 raw[256], srt[256], cur[256], nxt[256];
  __shared__ u32 s_cnt;
  Coin c = coin_load(ps);
  c.seed = r_merge_int(c.seed, ps->nonce); c.counter = 0;
  const u64 mask = (1ull << lN) - 1; const u32 lane = lane_id();
  for (u32 i = lane; i < num_queries; i += 32) { const Digest d = r_merge_int(c.seed, (u64)i + 1); raw[i] = (u32)(((u64)d.w[0] | ((u64)d.w[1] << 32)) & mask); }
  c.counter = num_queries;
  __syncwarp();
  for (u32 i = lane; i < num_queries; i += 32) {       // stable rank sort
    const u32 v = raw[i]; u32 r = 0;
    for (u32 j = 0; j < num_queries; j++) r += (raw[j] < v) || (raw[j] == v && j < i);
    srt[r] = v;
  }
  __syncwarp();
  // order-preserving compaction of `keep` flags: out[] gets the kept values, returns the count (warp-uniform)
  auto compact = [&](const u32* in, u32 n, u32 vmask, bool sorted_input, u32* out) -> u32 {
    u32 base = 0;
    for (u32 i0 = 0; i0 < n; i0 += 32) {
      const u32 i = i0 + lane; bool keep = false; u32 v = 0;
      if (i < n) {
        v = in[i] & vmask;
        if (sorted_input) keep = (i == 0) || ((in[i - 1] & vmask) != v);
        else { keep = true; for (u32 j = 0; j < i; j++) if ((in[j] & vmask) == v) { keep = false; break; } }
      }
      const u32 m = __ballot_sync(0xFFFFFFFFu, keep);
      if (keep) out[base + __popc(m & ((1u << lane) - 1))] = v;
      base += __popc(m);
    }
    __syncwarp();
    return base;
  };
  u32 cnt = compact(srt, num_queries, 0xFFFFFFFFu, true, cur);
  for (u32 i = lane; i < cnt; i += 32) ps->positions[i] = cur[i];
  if (lane == 0) ps->num_positions = cnt;
  u32 lNl = lN; u32* a = cur; u32* b = nxt;
  for (u32 l = 0; l < num_layers; l++) {
    const u32 fc = compact(a, cnt, (1u << (lNl - 3)) - 1, false, b);       // fold_positions: p mod (Nl/8), first occurrence kept
    for (u32 i = lane; i < fc; i += 32) ps->fri_positions[l][i] = b[i];
    if (lane == 0) ps->fri_num_positions[l] = fc;
    cnt = fc; u32* t = a; a = b; b = t; lNl -= 3;
    __syncwarp();
  }
  (void)s_cnt;
  coin_store(ps, c);
}

// one block row per task: queried rows and, for every queried leaf, the sibling digest at each level
__global__ void __launch_bounds__(256) gather_kernel(GatherTasks tasks, const ProofState* __restrict__ ps, u64* __restrict__ out) {
  const GatherTask& t = tasks.t[blockIdx.y];
  const u32 cnt = t.fri_layer < 0 ? ps->num_positions : ps->fri_num_positions[t.fri_layer];
  const u32* pos = t.fri_layer < 0 ? ps->positions : ps->fri_positions[t.fri_layer];
  const u32 width = t.J * t.limbs;
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < cnt * width; e += gridDim.x * blockDim.x) {
    const u32 q = e / width, w = e % width, j = w / t.limbs, l = w % t.limbs;
    const u64 p = (u64)pos[q] + (u64)j * t.R;
    const u64 addr = t.coset_n ? (p & 7) * t.coset_n + (p >> 3) : p;
    out[t.rows_off + e] = t.src[(size_t)l * t.limb_stride + addr];
  }
  Digest* paths = reinterpret_cast<Digest*>(out + t.paths_off);
  for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < cnt * t.depth; e += gridDim.x * blockDim.x) {
    const u32 q = e / t.depth, lvl = e % t.depth;
    const u64 node = ((t.M + pos[q]) >> lvl) ^ 1;
    paths[e] = t.tree[node];
  }
}

void launch_ood_finish(cudaStream_t st, int D, ProofState* ps, const u64* partial, u32 nb) {
  if (D == 1) ood_finish_kernel<1><<<1, 32 * (XFG_TRACE_WIDTH + 1), 0, st>>>(ps, partial, nb); else ood_finish_kernel<2><<<1, 32 * (XFG_TRACE_WIDTH + 2), 0, st>>>(ps, partial, nb);
  XFG_LAUNCHED(1);
}
void launch_grind(cudaStream_t st, ProofState* ps, u32 grinding) {
  // small grinding factors need a handful of candidates; large ones use the whole chip
  const unsigned blocks = grinding <= 8 ? 4 : (grinding <= 16 ? 148 : 148 * 8);
  grind_kernel<<<blocks, 256, 0, st>>>(ps, grinding);
  XFG_LAUNCHED(1);
}
void launch_positions(cudaStream_t st, ProofState* ps, u32 num_queries, u32 lN, u32 num_layers) { positions_kernel<<<1, 32, 0, st>>>(ps, num_queries, lN, num_layers); XFG_LAUNCHED(1); }
void launch_gather(cudaStream_t st, const GatherTasks& tasks, const ProofState* ps, u64* out) {
  gather_kernel<<<dim3(4, tasks.count), 256, 0, st>>>(tasks, ps, out);
  XFG_LAUNCHED(1);
}

}  // namespace xfg
