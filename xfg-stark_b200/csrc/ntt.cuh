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
  // direct (one 8-byte load per element) twiddle tables of a four-step transform, see NttTables; they replace pre_*, it_*, post_*
  const u64* pre_row; const u64* it_tab; u64 it_tstride; const u64* post_tab; u64 post_tstride;
  u32 grp_fast;   // ntt_pass_r16: blockIdx.y = sub * groups + group instead of group * src_div + sub
  // three-pass transforms (n = n1 n2 n3, 2^24 points and up: 2^12-point tiles would run one 1024-thread CTA per SM and leave 32-byte runs):
  //   tile t -> (hi, lo) = (t >> tile_lo_log, t mod 2^tile_lo_log); first column of the tile = hi * in_hi_stride + lo * T (out_hi_stride for the store)
  //   row_tw: non-transposed store multiplies output row k by row_tw[hi * k] (the twiddle between the two in-place passes)
  //   swap_lo_log / swap_hi_log: the transposed store of pass A writes column c = lo + 2^swap_lo_log * hi at (hi + 2^swap_hi_log * lo) * L
  u32 tile_lo_log; u64 in_hi_stride, out_hi_stride; const u64* row_tw; u32 swap_lo_log, swap_hi_log;
  // input validation fused into the first load of a transform: if non-null, *canon_flag |= canon_bit when an input element is >= p
  // (the prover's check that every trace element is a canonical field element; saves a separate pass over the trace)
  u32* canon_flag; u32 canon_bit;
};

// device tables owned by a plan (one per trace length)
struct NttTables {
  const u64* tw_fwd; const u64* tw_inv;   // w_4096^(+-i), i < 2048
  PowTable wn_fwd, wn_inv;                // powers of w_n and w_n^-1 (exponents < n)
  // Optional full-size tables for the four-step transforms of this length (n = n1 * n2, pass A = n1 column transforms of length n2).
  // The NTT kernels are bound by the integer pipe, not by HBM, so trading a two-level power lookup + 64-bit modular multiplication
  // per element for one more 8-byte load (mostly L2 hits: the tables are shared by all columns) is a net win.
  //   d_it_inv   [n]            (j1, k2) at j1*n2 + k2:  w_n^-(j1*k2) * d_scale             inverse transforms with scale == d_scale
  //   d_it_coset [cosets][n]                              w_n^(j1*k2)  * base_c^j1           forward coset transforms (pre table == d_pre_id)
  //   d_pre_row  [cosets][n2]   row j2:                   base_c^(j2*n1)                     the rest of the coset pre-scale base_c^(j1 + n1*j2)
  //   d_post     [posts][n]     output j:                 post_base_c^j                      inverse coset transforms (post table == d_post_id)
  const u64* d_it_inv = nullptr; const u64* d_it_coset = nullptr; const u64* d_pre_row = nullptr; const u64* d_post = nullptr;
  const u64* d_pre_id = nullptr; const u64* d_post_id = nullptr; u64 d_scale = 0; u32 d_ln = 0;
  u32 d_l2 = 0;                                            // length (log2) of pass A the d_it_* / d_pre_row tables were built for
  const u64* d_rtw_fwd = nullptr; const u64* d_rtw_inv = nullptr;   // three-pass plans: w_65536^(+-e), e < 2^16
};
// words of device memory ntt_build_direct needs / fill the tables (synchronous on stream 0); direct tables exist for 16 <= ln <= NTT_DIRECT_MAX_LOG
static constexpr u32 NTT_THREE_PASS_MIN_LOG = 24;
static constexpr u32 NTT_DIRECT_MIN_LOG = 16, NTT_DIRECT_MAX_LOG = 24;   // 1.4 GB of tables at 2^24 (falls back to the lookups when the allocation fails)
// length (log2) of pass A: half of ln for the four-step plan; ln - 16 for the three-pass plan (2^(ln-16) x 2^8 x 2^8) used from 2^24 points up
static inline u32 ntt_pass_a_log(u32 ln) { return ln >= NTT_THREE_PASS_MIN_LOG ? ln - 16 : ln / 2; }
size_t ntt_direct_words(u32 ln, u32 cosets, u32 posts);
void ntt_build_direct(NttTables& tb, u32 ln, u64* storage, u64 scale, const u64* pre_lo, const u64* pre_hi, u32 pre_hi_stride, u32 cosets,
                      const u64* post_lo, const u64* post_hi, u32 post_hi_stride, u32 posts);

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
  u32* canon_flag; u32 canon_bit;           // see NttPass (applies to the pass that reads `src`)
};

void ntt_init(bool force);
void ntt_batch(cudaStream_t st, const NttTables& tb, const NttJob& job);

}  // namespace xfg
