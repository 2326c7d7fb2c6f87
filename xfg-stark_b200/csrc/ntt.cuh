// ntt.cuh — interface of the batched NTT (see ntt.cu).
#pragma once
#include <cuda_runtime.h>
#include "field.cuh"

namespace xfg {

static constexpr int NTT_THREADS = 256;
static constexpr u32 NTT_TW_LOG = 12;          // twiddle table holds w_4096^i, i < 2048
static constexpr u32 NTT_SINGLE_MAX_LOG = 11;
static constexpr int NTT_MAX_PEERS = 8;       // one 8-GPU box  // transforms up to 2^11 points run as one shared-memory pass

struct NttPass {
  const u64* src; u64* dst;
  u32 Llog, Tlog;
  u64 in_row_stride, out_row_stride;
  u64 src_tstride, dst_tstride; u32 src_div;
  u32 store_transposed;
  const u64* tw;
  const u64* pre_lo; const u64* pre_hi; u32 pre_hi_stride;
  const u64* it_lo; const u64* it_hi;
  const u64* post_lo; const u64* post_hi; u32 post_hi_stride; u32 post_div;
  u64 scale;
  u32 num_radix, radix_logs;   // ntt_pass_r16: Stockham pass radices, 4 bits each (log2), first pass in the low nibble
  // fused all-to-all (config 5): when peer_log != 0 the last store of output element m of transform t goes to
  // peer[m >> peer_log][t * 2^peer_log + (m mod 2^peer_log)] (peer pointers may be NVLink-mapped memory of other GPUs)
  u64* peer[NTT_MAX_PEERS]; u32 peer_log;
  u32 coset_map, dst_cosets, src_is_dst;
};

// device tables owned by a plan (one per trace length)
struct NttTables {
  const u64* tw_fwd; const u64* tw_inv;   // w_4096^(+-i), i < 2048
  PowTable wn_fwd, wn_inv;                // powers of w_n and w_n^-1 (exponents < n)
};

// transform t (0 <= t < batch) reads src + (t / src_div) * src_tstride and writes dst + t * dst_tstride;
// if pre_lo != null, input element j is first multiplied by base_c^j, c = t % src_div, table c at pre_lo + c*4096 / pre_hi + c*pre_hi_stride
struct NttJob {
  const u64* src; u64* dst;
  u32 ln, batch;
  u64 src_tstride, dst_tstride; u32 src_div;
  bool inverse; u64 scale;                // scale = 1 or n^-1
  const u64* pre_lo; const u64* pre_hi; u32 pre_hi_stride;
  // if post_lo != null, output element j is multiplied by base_c^j, c = t % post_div
  const u64* post_lo; const u64* post_hi; u32 post_hi_stride; u32 post_div;
  u64* peer[NTT_MAX_PEERS]; u32 peer_log;   // see NttPass
  u32 coset_map, dst_cosets;                // optional subset of cosets, see ntt_pass
};

void ntt_init(bool force);
void ntt_batch(cudaStream_t st, const NttTables& tb, const NttJob& job);

}  // namespace xfg
