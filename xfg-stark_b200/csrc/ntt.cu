// ntt.cu — batched Goldilocks NTT / inverse NTT for sm_100a: shared-memory-staged radix-2 passes, four-step for n > 2^11.
//
// Replaces winter-math 0.8.4 `fft::{interpolate_poly, evaluate_poly_with_offset, interpolate_poly_with_offset}` as used by
// `ColMatrix::interpolate_columns`, `RowMatrix::evaluate_polys_over` and `CompositionPoly::new`
// (reference hook src/burn_mint_air.rs:504-514).  Natural order in, natural order out.
//
// One kernel, `ntt_pass`, does a length-L transform on T interleaved columns held in shared memory:
//   * single pass (n = L <= 2^11): T = 1, contiguous load/store;
//   * four-step (n = n1*n2, j = j1 + n1*j2, k = k2 + n2*k1):
//       pass A: tile = T consecutive j1, rows j2 (stride n1) -> size-n2 NTT, times w_n^(j1*k2), stored transposed Y[j1*n2 + k2];
//       pass B: tile = T consecutive k2, rows j1 (stride n2) -> size-n1 NTT, stored in place X[k2 + n2*k1].
// Global accesses are runs of T*8 bytes (pass A load, pass B load/store) or fully contiguous rows (pass A store).
// The coset pre-scale p[j] * s^j of the LDE is fused into the pass-A / single-pass load; 1/n and the coset un-scale
// c[j] * s^-j of interpolate_poly_with_offset into the last store.
#include "ntt.cuh"
#include "launch.cuh"

namespace xfg {

__device__ __forceinline__ u32 bitrev(u32 x, u32 bits) { return __brev(x) >> (32 - bits); }

__global__ void __launch_bounds__(NTT_THREADS) ntt_pass(NttPass p) {
  extern __shared__ u64 smem[];
  const u32 L = 1u << p.Llog, T = 1u << p.Tlog, TP = T > 1 ? T + 1 : 1;   // padded row: conflict-free transposed reads
  u64* S = smem;                   // S[row * TP + col]
  u64* TW = smem + (size_t)L * TP; // w_L^i, i < L/2
  const u32 tid = threadIdx.x, tile = blockIdx.x, tr = blockIdx.y;
  const u64* src = p.src + (size_t)(tr / p.src_div) * p.src_tstride;
  u64* dst = p.dst + (size_t)tr * p.dst_tstride;
  const u32 coset = tr % p.src_div;

  for (u32 i = tid; i < L / 2; i += NTT_THREADS) TW[i] = p.tw[(size_t)i << (NTT_TW_LOG - p.Llog)];

  // ---- load (bit-reversed row placement), optional pre-scale by s^(global index) ----
  const u64 col0 = (u64)tile << p.Tlog;
  PowTable pre; pre.lo = p.pre_lo ? p.pre_lo + (size_t)coset * POW_LO : nullptr; pre.hi = p.pre_hi ? p.pre_hi + (size_t)coset * p.pre_hi_stride : nullptr;
  for (u32 e = tid; e < L * T; e += NTT_THREADS) {
    u32 c = e & (T - 1), r = e >> p.Tlog;
    u64 gi = (u64)r * p.in_row_stride + col0 + c;
    u64 v = src[gi];
    if (pre.lo) v = gl_mul(v, pow_lookup(pre, gi));
    S[bitrev(r, p.Llog) * TP + c] = v;
  }
  __syncthreads();

  // ---- radix-2 DIT stages ----
  for (u32 s = 0; s < p.Llog; s++) {
    const u32 half = 1u << s, tshift = p.Llog - s - 1;
    for (u32 b = tid; b < (L / 2) * T; b += NTT_THREADS) {
      u32 c = b & (T - 1), bf = b >> p.Tlog;
      u32 j = bf & (half - 1), i0 = ((bf >> s) << (s + 1)) | j;
      u64* pu = S + i0 * TP + c; u64* pv = pu + half * TP;
      u64 u = *pu, v = gl_mul(*pv, TW[j << tshift]);
      *pu = gl_add(u, v); *pv = gl_sub(u, v);
    }
    __syncthreads();
  }

  // ---- store ----
  if (p.store_transposed) {
    // Y[(col0 + c) * L + k] = S[k][c] * w_n^((col0 + c) * k)
    PowTable it; it.lo = p.it_lo; it.hi = p.it_hi;
    for (u32 e = tid; e < L * T; e += NTT_THREADS) {
      u32 k = e & (L - 1), c = e >> p.Llog;
      u64 v = S[k * TP + c];
      u64 ex = (col0 + c) * (u64)k;
      if (ex) v = gl_mul(v, pow_lookup(it, ex));
      dst[(col0 + c) * (u64)L + k] = v;
    }
  } else {
    PowTable post; post.lo = p.post_lo ? p.post_lo + (size_t)(tr % p.post_div) * POW_LO : nullptr;
    post.hi = p.post_hi ? p.post_hi + (size_t)(tr % p.post_div) * p.post_hi_stride : nullptr;
    for (u32 e = tid; e < L * T; e += NTT_THREADS) {
      u32 c = e & (T - 1), k = e >> p.Tlog;
      u64 v = S[k * TP + c];
      u64 go = (u64)k * p.out_row_stride + col0 + c;
      if (p.scale != 1) v = gl_mul(v, p.scale);
      if (post.lo) v = gl_mul(v, pow_lookup(post, go));
      dst[go] = v;
    }
  }
}

size_t ntt_pass_smem(u32 Llog, u32 Tlog) {
  size_t L = size_t(1) << Llog, T = size_t(1) << Tlog, TP = T > 1 ? T + 1 : 1;
  return (L * TP + L / 2) * sizeof(u64);
}

void ntt_init() {
  static bool done = false;
  if (done) return;
  cudaFuncSetAttribute(ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_pass_smem(12, 2));
  done = true;
}

// Enqueue a batch of `batch` length-2^ln transforms.  src != dst is required when ln > NTT_SINGLE_MAX_LOG.
void ntt_batch(cudaStream_t st, const NttTables& tb, const NttJob& job) {
  ntt_init();
  const u32 ln = job.ln;
  NttPass p{};
  p.tw = job.inverse ? tb.tw_inv : tb.tw_fwd;
  p.src_div = job.src_div ? job.src_div : 1;
  p.src_tstride = job.src_tstride; p.dst_tstride = job.dst_tstride;
  p.pre_lo = job.pre_lo; p.pre_hi = job.pre_hi; p.pre_hi_stride = job.pre_hi_stride;
  p.post_div = job.post_div ? job.post_div : 1;
  if (ln <= NTT_SINGLE_MAX_LOG) {
    p.src = job.src; p.dst = job.dst; p.Llog = ln; p.Tlog = 0; p.in_row_stride = 1; p.out_row_stride = 1;
    p.store_transposed = 0; p.scale = job.scale;
    p.post_lo = job.post_lo; p.post_hi = job.post_hi; p.post_hi_stride = job.post_hi_stride;
    ntt_pass<<<dim3(1, job.batch), NTT_THREADS, ntt_pass_smem(ln, 0), st>>>(p); XFG_LAUNCHED(1);
    return;
  }
  // four-step: n = n1 * n2 with n2 = 2^l2 (pass A length), n1 = 2^l1 (pass B length)
  const u32 l2 = ln / 2, l1 = ln - l2;
  const u32 Tlog = l1 > 11 ? 2 : 3;     // 2^12-point tiles only fit 4 columns
  // pass A
  p.src = job.src; p.dst = job.dst; p.Llog = l2; p.Tlog = Tlog; p.in_row_stride = u64(1) << l1; p.out_row_stride = 0;
  p.store_transposed = 1; p.scale = 1;
  p.it_lo = job.inverse ? tb.wn_inv.lo : tb.wn_fwd.lo; p.it_hi = job.inverse ? tb.wn_inv.hi : tb.wn_fwd.hi;
  ntt_pass<<<dim3((1u << l1) >> Tlog, job.batch), NTT_THREADS, ntt_pass_smem(l2, Tlog), st>>>(p); XFG_LAUNCHED(1);
  // pass B (in place on dst)
  p.src = job.dst; p.dst = job.dst; p.src_div = 1; p.src_tstride = job.dst_tstride;
  p.pre_lo = nullptr; p.pre_hi = nullptr;
  p.Llog = l1; p.in_row_stride = u64(1) << l2; p.out_row_stride = u64(1) << l2; p.store_transposed = 0; p.scale = job.scale;
  p.post_lo = job.post_lo; p.post_hi = job.post_hi; p.post_hi_stride = job.post_hi_stride;
  ntt_pass<<<dim3((1u << l2) >> Tlog, job.batch), NTT_THREADS, ntt_pass_smem(l1, Tlog), st>>>(p); XFG_LAUNCHED(1);
}

}  // namespace xfg
