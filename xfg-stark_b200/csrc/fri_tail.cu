// fri_tail.cu — the latency-bound end of the proof in ONE launch: every FRI layer with at most 2^13 evaluations (FRI_TAIL_MAX_LOG) (tree, commitment,
// alpha, fold, rows of the next layer), the remainder (coset interpolation, commitment), the proof-of-work nonce and the query positions.
//
// Replaces the tail of winter-fri 0.8.3 `FriProver::build_layers` (per layer: `hash_values` -> `MerkleTree::new` -> `commit_fri_layer` ->
// `draw_fri_alpha` -> `apply_drp`; then the remainder `interpolate_poly_with_offset` + `hash_elements`) and winter-prover's grinding loop +
// `ProverChannel::get_query_positions` + winter-fri `fold_positions`, as reached from `air.prove(trace)` (src/burn_mint_prover.rs:124;
// SURVEY.md A.5, A.10).  Before this kernel those steps were 10-13 dependent launches of one warp to a few CTAs each (half of a 2^16-row proof).
// One CTA of 1024 threads walks the phases with block barriers; the Fiat-Shamir steps run on warp 0.  All BLAKE3 work goes through ONE
// non-inlined, rolled compression (b3r: 7 iterations of one round + the message permutation, ~2.5 KB of SASS): this code runs once per proof,
// so every instruction is a cold instruction-cache miss and code size, not issue rate, sets its latency.
#include "stark_kernels.cuh"
#include "transcript.cuh"
#include <algorithm>
#include "b3_rolled.cuh"
#include "launch.cuh"

namespace xfg {

namespace {

template <int D> __device__ __forceinline__ Ext<D> ld_ext(const u64 (*p)[2], int i) { return Ext<D>(p[i][0], p[i][1]); }

constexpr int TAIL_MAX_THREADS = 1024;

// blockDim.x = 1024 for a single proof (lowest latency), 256 in batch mode (a 1024-thread CTA at 64 registers needs a completely empty SM,
// which a busy multi-stream pipeline rarely offers: the 1024-thread tail cost the 1024-proof batch 9 %)
template <int D>
__global__ void __launch_bounds__(TAIL_MAX_THREADS, 1) fri_tail_kernel(FriTailArgs a, ProofState* ps) {
  extern __shared__ u64 sm[];                        // tree nodes of the current layer / remainder interpolation (2^rem_log x D); then the remainder limbs
  const u32 TAIL_THREADS = blockDim.x;
#ifdef XFG_TAIL_CLOCKS
  int ck = 0;
#define TCK() do { if (threadIdx.x == 0 && ck < 48) { ps->dbg_clk[ck++] = clock64(); } } while (0)
#else
#define TCK() do {} while (0)
#endif
  TCK();
  __shared__ u32 raw[256], srt[256], cur[256];
  __shared__ unsigned long long s_nonce;
  const u32 tid = threadIdx.x, lane = tid & 31;
  // ---- FRI layers first .. num_layers-1 (A.10) ----
  for (u32 l = a.first_layer; l < a.num_layers; l++) {
    const u32 lNl = a.layer_log[l]; const size_t Nl = size_t(1) << lNl, R = Nl >> 3, Rn = R >> 3;
    Digest* tree = a.tree[l];                        // leaves tree[R + i] were hashed by the kernel that produced this layer
    Digest* snode = reinterpret_cast<Digest*>(sm);   // the inner nodes are kept in shared memory (heap index), global memory only receives a copy for the query phase
    for (size_t cnt = R >> 1; cnt >= 1; cnt >>= 1) { // level with `cnt` nodes: heap indices [cnt, 2 cnt)
      for (size_t i = tid; i < cnt; i += TAIL_THREADS) {
        const size_t k = cnt + i; const Digest* ch = (cnt == (R >> 1)) ? tree + 2 * k : snode + 2 * k;
        const Digest d = r_merge(load_digest(ch), load_digest(ch + 1));
        store_digest(snode + k, d); store_digest(tree + k, d);
      }
      __syncthreads();
    }
    TCK();
    if (tid < 32) {                                  // commit_fri_layer + draw_fri_alpha
      Coin c = coin_load(ps); const Digest root = load_digest(snode + 1); c.seed = r_merge(c.seed, root); c.counter = 0;
      const bool ok = r_draw_many<D>(c, 1, &ps->alphas[l]);
      if (lane == 0) { ps->fri_roots[l] = root; if (!ok) ps->error_flags |= ERR_FLAG_COIN; }
      coin_store(ps, c);
    }
    __syncthreads();
    TCK();
    // apply_drp: next[r] = P_r(alpha), P_r interpolating row r = values at r + j R over x_r w_8^j, x_r = 7 w_Nl^r
    const Ext<D> alpha = ld_ext<D>(ps->alphas, l);
    const u64* src = a.evals[l]; u64* dst = a.evals[l + 1];
    const size_t sstride = a.limb_stride[l], dstride = a.limb_stride[l + 1];
    const bool coset = l == 0;                       // layer 0 = DEEP evaluations, coset-major [k][m] for LDE row 8 m + k
    for (size_t r = tid; r < R; r += TAIL_THREADS) {
      Ext<D> v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const size_t p = r + (size_t)j * R, addr = coset ? (p & 7) * (Nl >> 3) + (p >> 3) : p;
#pragma unroll
        for (int q = 0; q < D; q++) v[j].set_limb(q, src[(size_t)q * sstride + addr]);
      }
      const u64 xinv = gl_mul(a.fc.inv7, pow_lookup(a.wN_inv, (u64)r << (a.lN - lNl)));
      const Ext<D> w = fold8<D>(v, a.fc, mul_base(alpha, xinv));
#pragma unroll
      for (int q = 0; q < D; q++) dst[(size_t)q * dstride + r] = w.limb(q);
    }
    __syncthreads();
    TCK();
    if (l + 1 < a.num_layers) {                      // hash_values: leaves of the next layer's tree, row i' = values at i' + q Rn
      Digest* nt = a.tree[l + 1];
      for (size_t ip = tid; ip < Rn; ip += TAIL_THREADS) {
        u64 row[8 * D];
#pragma unroll
        for (int q = 0; q < 8; q++)
#pragma unroll
          for (int k = 0; k < D; k++) row[q * D + k] = dst[(size_t)k * dstride + ip + (size_t)q * Rn];
        store_digest(nt + Rn + ip, r_chunk(row, 8 * D, 0, true));
      }
      __syncthreads();
    }
    TCK();
  }
  // ---- remainder: coset interpolation (offset 7) of the last layer's evaluations, first rem_len coefficients (A.10) ----
  {
    const u32 Llog = a.rem_log, L = 1u << Llog;
    for (u32 e = tid; e < L * D; e += TAIL_THREADS) { const u32 q = e >> Llog, i = e & (L - 1); sm[q * L + (__brev(i) >> (32 - Llog))] = a.rem_in[(size_t)q * a.rem_stride + i]; }
    __syncthreads();
    for (u32 s = 0; s < Llog; s++) {
      const u32 half = 1u << s, tshift = NTT_TW_LOG_TAIL - s - 1;
      for (u32 b = tid; b < (L / 2) * D; b += TAIL_THREADS) {
        const u32 q = b / (L / 2), bf = b % (L / 2), j = bf & (half - 1), i0 = ((bf >> s) << (s + 1)) | j;
        u64* pu = sm + q * L + i0; u64* pv = pu + half;
        const u64 u = *pu, v = gl_mul(*pv, a.tw_inv[(size_t)j << tshift]);
        *pu = gl_add(u, v); *pv = gl_sub(u, v);
      }
      __syncthreads();
    }
    TCK();
    // c_j = X_j / L * 7^-j ; limbs interleaved per coefficient for hash_elements
    u64* limbs = sm + a.limbs_off;
    for (u32 e = tid; e < a.rem_len * 2; e += TAIL_THREADS) {
      const u32 i = e >> 1, q = e & 1;
      u64 v = 0;
      if (q < (u32)D) { v = gl_mul(gl_mul(sm[q * L + i], a.rem_ninv), a.un_lo[i]); limbs[i * D + q] = v; }
      ps->remainder[i][q] = v;
    }
    __syncthreads();
    if (tid < 32) {
      Coin c = coin_load(ps);
      const Digest d = r_hash_limbs(limbs, (int)(a.rem_len * D));       // every lane computes the same digest (keeps the coin warp-uniform)
      c.seed = r_merge(c.seed, d); c.counter = 0;
      if (lane == 0) { ps->remainder_len = a.rem_len; ps->remainder_commitment = d; }
      coin_store(ps, c);
    }
    if (tid == 0) s_nonce = ~0ull;
    __syncthreads();
    TCK();
  }
  if (!a.do_grind) return;                           // large grinding factors: the multi-CTA grind kernel and positions_kernel follow
  // ---- grinding: smallest nonce >= 1 whose hash has `grinding` trailing zero bits (A.5); candidates in rounds of 1024 ----
  {
    const Digest seed = ps->seed;
    const u64 mask = a.grinding >= 64 ? ~0ull : ((1ull << a.grinding) - 1);
    for (u64 base = 1;; base += TAIL_THREADS) {   // rounds of blockDim.x candidates: the minimum over a round is the serial minimum
      const u64 nonce = base + tid;
      const Digest d = r_merge_int(seed, nonce);
      if ((((u64)d.w[0] | ((u64)d.w[1] << 32)) & mask) == 0) atomicMin(&s_nonce, (unsigned long long)nonce);
      __syncthreads();
      if (s_nonce != ~0ull) break;                   // uniform: every thread reads the same value after the barrier
      __syncthreads();
    }
    if (tid == 0) ps->nonce = s_nonce;
    TCK();
  }
  // ---- draw_integers(q, N, nonce) -> sort -> dedup (warp 0), then fold_positions per layer (A.5, A.10).  The folded list of layer l is the
  // order-preserving dedup of (previous list mod N_l/8); a value dropped at an earlier layer is a duplicate at every later one too, so it equals
  // the order-preserving dedup of (sorted unique positions mod N_l/8) and the layers are independent: warp l folds layer l.
  __shared__ u32 s_cnt;
  const u32 num_queries = a.num_queries; const u64 pmask = (1ull << a.lN) - 1;
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
  if (tid < 32) {
    Coin c = coin_load(ps);
    c.seed = r_merge_int(c.seed, s_nonce); c.counter = 0;
    for (u32 i = lane; i < num_queries; i += 32) { const Digest d = r_merge_int(c.seed, (u64)i + 1); raw[i] = (u32)(((u64)d.w[0] | ((u64)d.w[1] << 32)) & pmask); }
    c.counter = num_queries;
    __syncwarp();
    for (u32 i = lane; i < num_queries; i += 32) {       // stable rank sort
      const u32 v = raw[i]; u32 r = 0;
      for (u32 j = 0; j < num_queries; j++) r += (raw[j] < v) || (raw[j] == v && j < i);
      srt[r] = v;
    }
    __syncwarp();
    const u32 cnt = compact(srt, num_queries, 0xFFFFFFFFu, true, cur);
    for (u32 i = lane; i < cnt; i += 32) ps->positions[i] = cur[i];
    if (lane == 0) { ps->num_positions = cnt; s_cnt = cnt; }
    coin_store(ps, c);
  }
  __syncthreads();
  const u32 nwarps = TAIL_THREADS >> 5, warp = tid >> 5;
  for (u32 l = warp; l < a.num_layers; l += nwarps) {
    u32* outp = ps->fri_positions[l];                      // written in place: lanes only read `cur`
    const u32 fcn = compact(cur, s_cnt, (1u << (a.lN - 3 * (l + 1))) - 1, false, outp);
    if (lane == 0) ps->fri_num_positions[l] = fcn;
  }
  TCK();
}

}  // namespace

void launch_fri_tail(cudaStream_t st, int D, FriTailArgs a, ProofState* ps, u32 threads) {
  static thread_local bool attr = false;
  constexpr size_t MAX_SMEM = (size_t(32) << (FRI_TAIL_MAX_LOG - 3)) + (size_t)MAX_REMAINDER * 2 * 8;   // 2^11 nodes of 32 bytes + remainder limbs
  if (!attr) { cudaFuncSetAttribute(fri_tail_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_SMEM);
               cudaFuncSetAttribute(fri_tail_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_SMEM); attr = true; }
  size_t words = (size_t)D << a.rem_log;                                              // remainder interpolation
  for (u32 l = a.first_layer; l < a.num_layers; l++) words = std::max(words, (size_t(4) << (a.layer_log[l] - 3)));   // R digests = 4 R words
  a.limbs_off = (u32)words;
  const size_t smem = (words + (size_t)MAX_REMAINDER * 2) * 8;
  if (D == 1) fri_tail_kernel<1><<<1, threads, smem, st>>>(a, ps); else fri_tail_kernel<2><<<1, threads, smem, st>>>(a, ps);
  XFG_LAUNCHED(1);
}

}  // namespace xfg
