// transcript.cu — the Fiat-Shamir channel on the device: coin seeding / reseeding / draws, grinding, query positions and
// the gather of queried rows and authentication nodes.
//
// Replaces winter-prover 0.8.3 `ProverChannel::{new, commit_trace, commit_constraints, get_*_coeffs, get_ood_point,
// send_ood_*, commit_fri_layer, draw_fri_alpha, grind_query_seed, get_query_positions}`, winter-crypto
// `DefaultRandomCoin`, winter-fri `fold_positions` and the row/path collection of `DefaultTraceLde::query`,
// `ConstraintCommitment::query`, `FriProver::build_proof` (A.4, A.5, A.9-A.11).
// The serial steps are single-thread kernels (a handful of BLAKE3 compressions each) so the proof needs no host round trip.
#include "transcript.cuh"
#include "launch.cuh"

namespace xfg {

struct Coin {
  Digest seed; u64 counter;
  __device__ void reseed(const Digest& d) { seed = b3_merge(seed, d); counter = 0; }
  __device__ Digest next() { counter += 1; return b3_merge_int(seed, counter); }
  // draw::<E>(): first 8*D bytes of next(); every limb must be canonical, else retry (A.5)
  template <int D> __device__ bool draw(u64 out[2]) {
    for (int t = 0; t < XFG_COIN_MAX_DRAWS; t++) {
      Digest d = next();
      u64 v0 = (u64)d.w[0] | ((u64)d.w[1] << 32), v1 = (u64)d.w[2] | ((u64)d.w[3] << 32);
      if (v0 < GL_P && (D == 1 || v1 < GL_P)) { out[0] = v0; out[1] = D == 2 ? v1 : 0; return true; }
    }
    out[0] = out[1] = 0; return false;
  }
};
__device__ __forceinline__ Coin coin_load(const ProofState* ps) { Coin c; c.seed = ps->seed; c.counter = ps->counter; return c; }
__device__ __forceinline__ void coin_store(ProofState* ps, const Coin& c) { ps->seed = c.seed; ps->counter = c.counter; }
template <int D> __device__ __forceinline__ Ext<D> ldx(const u64* p) { return Ext<D>(p[0], p[1]); }
template <int D> __device__ __forceinline__ void stx(u64* p, Ext<D> v) { p[0] = v.limb(0); p[1] = D == 2 ? v.limb(1) : 0; }

// coin = hash_elements(context elements || public inputs)  (A.4)
__global__ void seed_kernel(ProofState* ps, const u64* __restrict__ seed_limbs, int count) {
  ps->seed = b3_hash_limbs_dyn(seed_limbs, count); ps->counter = 0; ps->error_flags = 0; ps->nonce = ~0ull;
}
template <int D> __global__ void trace_root_kernel(ProofState* ps, const Digest* __restrict__ tree) {
  Coin c = coin_load(ps); Digest root = tree[1]; ps->trace_root = root; c.reseed(root);
  bool ok = true;
  for (int j = 0; j < XFG_NUM_TRANSITION; j++) ok &= c.draw<D>(ps->tcoef[j]);   // transition coefficients first, then boundary (A.8)
  for (int j = 0; j < XFG_NUM_ASSERTIONS; j++) ok &= c.draw<D>(ps->bcoef[j]);
  if (!ok) ps->error_flags |= ERR_FLAG_COIN;
  coin_store(ps, c);
}
template <int D> __global__ void constraint_root_kernel(ProofState* ps, const Digest* __restrict__ tree, u64 g_n) {
  Coin c = coin_load(ps); Digest root = tree[1]; ps->constraint_root = root; c.reseed(root);
  if (!c.draw<D>(ps->z)) ps->error_flags |= ERR_FLAG_COIN;
  stx<D>(ps->zg, mul_base(ldx<D>(ps->z), g_n));
  coin_store(ps, c);
}
// sums the OOD partials, sends the frame and H(z) to the coin, draws the DEEP coefficients (A.9)
template <int D> __global__ void ood_finish_kernel(ProofState* ps, const u64* __restrict__ partial, u32 nb) {
  Coin c = coin_load(ps);
  u64 sums[NUM_OOD_POLYS][2][2];
  for (int p = 0; p < XFG_TRACE_WIDTH + D; p++)
    for (int w = 0; w < 2; w++) for (int l = 0; l < 2; l++) {
      u64 s = 0; for (u32 b = 0; b < nb; b++) s = gl_add(s, partial[(((size_t)p * nb + b) * 2 + w) * 2 + l]);
      sums[p][w][l] = s;
    }
  u64 limbs[2 * XFG_TRACE_WIDTH * 2]; int k = 0;
  for (int j = 0; j < XFG_TRACE_WIDTH; j++) for (int w = 0; w < 2; w++) {      // interleaved per column (A.9, D)
    ps->ood_frame[2 * j + w][0] = sums[j][w][0]; ps->ood_frame[2 * j + w][1] = D == 2 ? sums[j][w][1] : 0;
    for (int l = 0; l < D; l++) limbs[k++] = sums[j][w][l];
  }
  c.reseed(b3_hash_limbs_dyn(limbs, k));
  // H(z) = P_limb0(z) + phi * P_limb1(z), phi = (0,1): (a0,a1) * phi = (-2 a1, a0 + a1)
  Ext<D> hz = ldx<D>(sums[XFG_TRACE_WIDTH][0]);
  if (D == 2) { u64 a0 = sums[XFG_TRACE_WIDTH + 1][0][0], a1 = sums[XFG_TRACE_WIDTH + 1][0][1]; hz = hz + Ext<D>(gl_neg(gl_dbl(a1)), gl_add(a0, a1)); }
  stx<D>(ps->hz, hz);
  u64 hl[2] = {hz.limb(0), hz.limb(1)};
  c.reseed(b3_hash_limbs_dyn(hl, D));
  bool ok = true;
  for (int j = 0; j <= XFG_TRACE_WIDTH; j++) ok &= c.draw<D>(ps->dcoef[j]);   // 7 trace coefficients, then 1 composition column
  if (!ok) ps->error_flags |= ERR_FLAG_COIN;
  Ext<D> c1, c2;
  for (int j = 0; j < XFG_TRACE_WIDTH; j++) {
    Ext<D> g = ldx<D>(ps->dcoef[j]);
    c1 = c1 + g * ldx<D>(ps->ood_frame[2 * j]); c2 = c2 + g * ldx<D>(ps->ood_frame[2 * j + 1]);
  }
  c1 = c1 + ldx<D>(ps->dcoef[XFG_TRACE_WIDTH]) * hz;
  stx<D>(ps->deep_c1, c1); stx<D>(ps->deep_c2, c2);
  coin_store(ps, c);
}
template <int D> __global__ void fri_commit_kernel(ProofState* ps, const Digest* __restrict__ tree, u32 layer) {
  Coin c = coin_load(ps); Digest root = tree[1]; ps->fri_roots[layer] = root; c.reseed(root);
  if (!c.draw<D>(ps->alphas[layer])) ps->error_flags |= ERR_FLAG_COIN;
  coin_store(ps, c);
}
// remainder = first `len` coefficients; commitment = hash_elements(remainder); reseed (A.10)
template <int D> __global__ void remainder_kernel(ProofState* ps, const u64* __restrict__ coef, size_t limb_stride, u32 len) {
  Coin c = coin_load(ps);
  u64 limbs[MAX_REMAINDER * 2];
  for (u32 i = 0; i < len; i++) for (int l = 0; l < 2; l++) {
    u64 v = l < D ? coef[(size_t)l * limb_stride + i] : 0;
    ps->remainder[i][l] = v; if (l < D) limbs[i * D + l] = v;
  }
  ps->remainder_len = len;
  Digest d = b3_hash_limbs_dyn(limbs, len * D);
  ps->remainder_commitment = d; c.reseed(d);
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
// draw_integers(q, N, nonce) -> sort -> dedup; then fold_positions per FRI layer (A.5, A.10)
__global__ void positions_kernel(ProofState* ps, u32 num_queries, u32 lN, u32 num_layers) {
  Coin c = coin_load(ps);
  c.seed = b3_merge_int(c.seed, ps->nonce); c.counter = 0;
  const u64 mask = (1ull << lN) - 1;
  u32 pos[MAX_Q];
  for (u32 i = 0; i < num_queries; i++) { Digest d = c.next(); pos[i] = (u32)(((u64)d.w[0] | ((u64)d.w[1] << 32)) & mask); }
  for (u32 i = 1; i < num_queries; i++) { u32 v = pos[i]; int j = (int)i - 1; while (j >= 0 && pos[j] > v) { pos[j + 1] = pos[j]; j--; } pos[j + 1] = v; }
  u32 cnt = 0;
  for (u32 i = 0; i < num_queries; i++) if (i == 0 || pos[i] != pos[i - 1]) pos[cnt++] = pos[i];
  ps->num_positions = cnt;
  for (u32 i = 0; i < cnt; i++) ps->positions[i] = pos[i];
  u32 lNl = lN;
  for (u32 l = 0; l < num_layers; l++) {
    const u32 tmask = (1u << (lNl - 3)) - 1; u32 fc = 0;
    for (u32 i = 0; i < cnt; i++) {                    // order-preserving dedup, not re-sorted
      u32 q = pos[i] & tmask; bool dup = false;
      for (u32 j = 0; j < fc; j++) if (ps->fri_positions[l][j] == q) { dup = true; break; }
      if (!dup) ps->fri_positions[l][fc++] = q;
    }
    ps->fri_num_positions[l] = fc;
    cnt = fc; for (u32 i = 0; i < cnt; i++) pos[i] = ps->fri_positions[l][i];
    lNl -= 3;
  }
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

void launch_seed(cudaStream_t st, ProofState* ps, const u64* seed_limbs, int count) { seed_kernel<<<1, 1, 0, st>>>(ps, seed_limbs, count); XFG_LAUNCHED(1); }
void launch_trace_root(cudaStream_t st, int D, ProofState* ps, const Digest* tree) {
  if (D == 1) trace_root_kernel<1><<<1, 1, 0, st>>>(ps, tree); else trace_root_kernel<2><<<1, 1, 0, st>>>(ps, tree);
  XFG_LAUNCHED(1);
}
void launch_constraint_root(cudaStream_t st, int D, ProofState* ps, const Digest* tree, u64 g_n) {
  if (D == 1) constraint_root_kernel<1><<<1, 1, 0, st>>>(ps, tree, g_n); else constraint_root_kernel<2><<<1, 1, 0, st>>>(ps, tree, g_n);
  XFG_LAUNCHED(1);
}
void launch_ood_finish(cudaStream_t st, int D, ProofState* ps, const u64* partial, u32 nb) {
  if (D == 1) ood_finish_kernel<1><<<1, 1, 0, st>>>(ps, partial, nb); else ood_finish_kernel<2><<<1, 1, 0, st>>>(ps, partial, nb);
  XFG_LAUNCHED(1);
}
void launch_fri_commit(cudaStream_t st, int D, ProofState* ps, const Digest* tree, u32 layer) {
  if (D == 1) fri_commit_kernel<1><<<1, 1, 0, st>>>(ps, tree, layer); else fri_commit_kernel<2><<<1, 1, 0, st>>>(ps, tree, layer);
  XFG_LAUNCHED(1);
}
void launch_remainder(cudaStream_t st, int D, ProofState* ps, const u64* coef, size_t limb_stride, u32 len) {
  if (D == 1) remainder_kernel<1><<<1, 1, 0, st>>>(ps, coef, limb_stride, len); else remainder_kernel<2><<<1, 1, 0, st>>>(ps, coef, limb_stride, len);
  XFG_LAUNCHED(1);
}
void launch_grind(cudaStream_t st, ProofState* ps, u32 grinding) {
  // small grinding factors need a handful of candidates; large ones use the whole chip
  const unsigned blocks = grinding <= 8 ? 4 : (grinding <= 16 ? 148 : 148 * 8);
  grind_kernel<<<blocks, 256, 0, st>>>(ps, grinding);
  XFG_LAUNCHED(1);
}
void launch_positions(cudaStream_t st, ProofState* ps, u32 num_queries, u32 lN, u32 num_layers) { positions_kernel<<<1, 1, 0, st>>>(ps, num_queries, lN, num_layers); XFG_LAUNCHED(1); }
void launch_gather(cudaStream_t st, const GatherTasks& tasks, const ProofState* ps, u64* out) {
  gather_kernel<<<dim3(4, tasks.count), 256, 0, st>>>(tasks, ps, out);
  XFG_LAUNCHED(1);
}

}  // namespace xfg
