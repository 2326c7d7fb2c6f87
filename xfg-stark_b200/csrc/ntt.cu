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
#include "field_weak.cuh"
#include "launch.cuh"

namespace xfg {

__device__ __forceinline__ u32 bitrev(u32 x, u32 bits) { return __brev(x) >> (32 - bits); }

__global__ void __launch_bounds__(NTT_THREADS) ntt_pass(NttPass p) {
  extern __shared__ u64 smem[];
  const u32 L = 1u << p.Llog, T = 1u << p.Tlog, TP = T > 1 ? T + 1 : 1;   // padded row: conflict-free transposed reads
  u64* S = smem;                   // S[row * TP + col]
  u64* TW = smem + (size_t)L * TP; // w_L^i, i < L/2
  const u32 tid = threadIdx.x, tile = blockIdx.x, tr = blockIdx.y;
  // transform tr = (group, sub): group selects the source polynomial, sub the coset.  coset_map (4 bits per entry, 0 = identity)
  // lets a job compute a subset of the cosets: table / output slot of sub is map[sub], the output has dst_cosets slots per group
  const u32 grp = tr / p.src_div, sub = tr % p.src_div;
  const u32 coset = p.coset_map ? (u32)((p.coset_map >> (4 * sub)) & 15) : sub;
  u64* dst = p.dst + (p.coset_map ? (size_t)grp * p.dst_cosets + coset : (size_t)tr) * p.dst_tstride;
  const u64* src = p.src_is_dst ? dst : p.src + (size_t)grp * p.src_tstride;

  for (u32 i = tid; i < L / 2; i += NTT_THREADS) TW[i] = p.tw[(size_t)i << (NTT_TW_LOG - p.Llog)];

  // ---- load (bit-reversed row placement), optional pre-scale by s^(global index) ----
  const u64 col0 = (u64)tile << p.Tlog;
  PowTable pre; pre.lo = p.pre_lo ? p.pre_lo + (size_t)coset * POW_LO : nullptr; pre.hi = p.pre_hi ? p.pre_hi + (size_t)coset * p.pre_hi_stride : nullptr;
  bool bad = false;
  for (u32 e = tid; e < L * T; e += NTT_THREADS) {
    u32 c = e & (T - 1), r = e >> p.Tlog;
    u64 gi = (u64)r * p.in_row_stride + col0 + c;
    u64 v = src[gi];
    bad |= v >= GL_P;
    if (pre.lo) v = gl_mul(v, pow_lookup(pre, gi));
    S[bitrev(r, p.Llog) * TP + c] = v;
  }
  if (p.canon_flag && bad) atomicOr(p.canon_flag, p.canon_bit);
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
      if (p.peer_log) p.peer[go >> p.peer_log][(size_t)tr * (u64(1) << p.peer_log) + (go & ((u64(1) << p.peer_log) - 1))] = v;   // row owner = m / n_local
      else dst[go] = v;
    }
  }
}


// =====================================================================================================================
// ntt_pass_r16 — the fast path for tiles of 2^8 .. 2^12 points: Stockham auto-sort passes of radix 16 / 8 / 4 held in
// registers.  Each thread owns 16 elements per pass (one radix-16 item, two radix-8 items or four radix-4 items), so a pass
// is: load 16 values from shared memory, multiply by the inter-pass twiddles w_{Ns*R}^(k*r), run the in-register DFTs,
// barrier, store, barrier.  Inside a radix-R DFT (R <= 16) every twiddle is a power of two (w_16 = 2^12, w_8 = 2^24,
// w_4 = 2^48, w_2 = 2^96 = -1), i.e. shifts and carry fix-ups on weak values instead of 64x64-bit multiplications; a
// 1024-point tile needs 1.7 full multiplications per element instead of 5.  Same load/store stages as ntt_pass.
// =====================================================================================================================
__host__ __device__ constexpr int brev_c(int x, int bits) { int r = 0; for (int i = 0; i < bits; i++) if (x >> i & 1) r |= 1 << (bits - 1 - i); return r; }

// (a, b) <- (a + w b, a - w b) with w = 2^S (forward) or w = 2^-S = -2^(96-S) (inverse); a, b weak in and out
template <int S, bool INV> __device__ __forceinline__ void bfly_pow2(u64& a, u64& b) {
  if (S == 0) { const u64 t = w_canon(b); const u64 x = w_add_c(a, t), y = w_sub_c(a, t); a = x; b = y; }
  else if (!INV) { const u64 t = w_canon(w_mul_pow2<S>(b)); const u64 x = w_add_c(a, t), y = w_sub_c(a, t); a = x; b = y; }
  else { const u64 t = w_canon(w_mul_pow2<(96 - S) % 96>(b)); const u64 x = w_sub_c(a, t), y = w_add_c(a, t); a = x; b = y; }
}
// radix-2 DIT over registers; element i of the (bit-reversed) working order lives in v[brev(i)], so the natural-order
// output k ends up in v[brev(k)] and no register is ever moved
template <int LOGR, bool INV, int STAGE, int IDX> struct DftStep {
  static __device__ __forceinline__ void run(u64 (&v)[1 << LOGR]) {
    constexpr int half = 1 << STAGE, j = IDX & (half - 1), g = IDX >> STAGE, i0 = g * 2 * half + j, i1 = i0 + half;
    bfly_pow2<j * (96 / half), INV>(v[brev_c(i0, LOGR)], v[brev_c(i1, LOGR)]);
    DftStep<LOGR, INV, STAGE, IDX + 1>::run(v);
  }
};
template <int LOGR, bool INV, int STAGE> struct DftStep<LOGR, INV, STAGE, (1 << LOGR) / 2> {
  static __device__ __forceinline__ void run(u64 (&v)[1 << LOGR]) { DftStep<LOGR, INV, STAGE + 1, 0>::run(v); }
};
template <int LOGR, bool INV> struct DftStep<LOGR, INV, LOGR, 0> { static __device__ __forceinline__ void run(u64 (&)[1 << LOGR]) {} };

// Inter-pass twiddles of a pass whose sub-transform length ns * R is 64: w_64 = 2^3, so the twiddle w_64^(+-k r) of input r is the
// compile-time shift 2^(+-3 K r) once k = K is known - one IMAD.WIDE and 8 - 13 instructions instead of a full 64 x 64-bit product
template <int K, bool INV, int R, int r = 1> struct Pow2Twiddle {
  static __device__ __forceinline__ void run(u64 (&v)[R]) {
    constexpr int e = (3 * K * r) % 192, E = INV ? (192 - e) % 192 : e;
    v[r] = w_mul_pow2_any<E>(v[r]);
    Pow2Twiddle<K, INV, R, r + 1>::run(v);
  }
};
template <int K, bool INV, int R> struct Pow2Twiddle<K, INV, R, R> { static __device__ __forceinline__ void run(u64 (&)[R]) {} };

// POW2 (ns = 4, R = 16, one item per thread): the items are dealt to the threads so that k = j mod 4 is the same for a whole warp
// (bits [1:0] of the item's row swapped with the two lowest warp-uniform bits), and the warp branches once on k into the shift twiddles.
template <int LOGR, bool INV, int EPT, bool POW2 = false>
__device__ __forceinline__ void stockham_pass(u64* __restrict__ S, const u64* __restrict__ TW, u32 Llog, u32 Tlog, u32 TP, u32 ns_log, u32 tid, u32 nthreads) {
  constexpr int R = 1 << LOGR, ITEMS = EPT / R;
  static_assert(!POW2 || (LOGR == 4 && ITEMS == 1), "POW2: radix 16, one item per thread");
  const u32 Tm = (1u << Tlog) - 1, stride = 1u << (Llog - LOGR), nsm = (1u << ns_log) - 1, tsh = Llog - ns_log - LOGR;
  const u32 wl = 5 - Tlog;      // POW2: rows per warp = 2^wl (>= 4: Tlog <= 3)
  auto row_of = [&](u32 jt) -> u32 { return POW2 ? (jt & ~(3u | (3u << wl))) | ((jt & 3u) << wl) | ((jt >> wl) & 3u) : jt; };
  u64 v[ITEMS][R];
#pragma unroll
  for (int q = 0; q < ITEMS; q++) {
    const u32 w = tid + q * nthreads, col = w & Tm, j = row_of(w >> Tlog), k = j & nsm;
#pragma unroll
    for (int r = 0; r < R; r++) v[q][r] = S[(j + r * stride) * TP + col];
    if (POW2) {
      switch (k) {      // warp-uniform
        case 1: Pow2Twiddle<1, INV, R>::run(v[q]); break;
        case 2: Pow2Twiddle<2, INV, R>::run(v[q]); break;
        case 3: Pow2Twiddle<3, INV, R>::run(v[q]); break;
        default: break;
      }
    } else if (ns_log) {
#pragma unroll
      for (int r = 1; r < R; r++) v[q][r] = w_mul(v[q][r], TW[(k * r) << tsh]);
    }
    DftStep<LOGR, INV, 0, 0>::run(v[q]);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < ITEMS; q++) {
    const u32 w = tid + q * nthreads, col = w & Tm, j = row_of(w >> Tlog), k = j & nsm;
    const u32 j0 = ((j >> ns_log) << (ns_log + LOGR)) | k;
#pragma unroll
    for (int r = 0; r < R; r++) S[(j0 + ((u32)r << ns_log)) * TP + col] = v[q][brev_c(r, LOGR)];
  }
  __syncthreads();
}

// tile shape and radix plan of ntt_pass_r16 as functions of the tile length (shared by the host launcher and the specialised kernels):
// 4096 elements per CTA (more with 2^11, 2^12-point tiles), 16 per thread: 8:(4,4) 9:(4,3,2) 10:(4,4,2) 11:(4,4,3) 12:(4,4,4)
#ifndef XFG_R16_T10LOG
#define XFG_R16_T10LOG 3      // log2 columns of a 2^10-point tile: 3 = 8 columns (64-byte runs), 8192 elements, 512 threads, 2 CTAs/SM; 2 = 4 columns, 256 threads, 4 CTAs/SM.
                              // Measured at 2^20 rows (round 2): trace LDE 1.165 -> 1.126 ms with 8 columns, the 4-transform composition interpolation 0.099 -> 0.107
#endif
__host__ __device__ constexpr u32 r16_tlog_c(u32 Llog) { return Llog == 10 ? XFG_R16_T10LOG : Llog >= 10 ? 2 : 12 - Llog; }
__host__ __device__ constexpr u32 r16_radix_count_c(u32 Llog) { return Llog == 8 ? 2 : 3; }
__host__ __device__ constexpr u32 r16_radix_packed_c(u32 Llog) {
  return Llog == 8 ? 0x44u : Llog == 9 ? 0x234u : Llog == 10 ? 0x244u : Llog == 11 ? 0x344u : 0x444u;
}

// XFG_R16_POW2TW (2^10-point tiles): 0 = radix plan 16 x 16 x 4, every inter-pass twiddle from the table; 1 = 4 x 16 x 16 with shift twiddles
// between the first two passes (Pow2Twiddle: 0.94 full multiplications per element instead of 1.69); 2 = the 4 x 16 x 16 order with table twiddles.
// Measured at 2^20 rows / quadratic (B200, round 2; parity green with 1): trace LDE 1.124 (0) / 1.121 (1) / 1.186 ms (2), the four NTT families together
// 1.664 / 1.668 / 1.750 ms.  The shift form takes three of a product's four IMAD.WIDE off the heavy FMA half, but costs as many ALU-pipe
// instructions (shifts and carry fix-ups: 10 - 11 against ~10), and the ALU pipe is the one that binds: neutral, so the simpler plan stays.
#ifndef XFG_R16_POW2TW
#define XFG_R16_POW2TW 0
#endif
#ifndef XFG_R16_MINB
#define XFG_R16_MINB 2   // 64 registers: measured best (1: 126 regs 2.26 ms, 2: 1.77 ms, 3: 1.84 ms, 4: 1.95 ms for the 2^20 trace LDE)
#endif
// LLOG != 0 specialises the kernel for 2^LLOG-point tiles of the standard shape: tile length, column count, thread count and the
// radix plan become compile-time constants, which turns the shared-memory index arithmetic of every pass (a fifth of the
// instructions of the generic kernel) into immediate offsets.  LLOG = 0 is the generic kernel (any shape the launcher passes).
template <bool INV, int EPT, int MAXT, int LLOG>
__global__ void __launch_bounds__(MAXT, MAXT == 1024 ? 1 : XFG_R16_MINB) ntt_pass_r16(NttPass p) {
  extern __shared__ u64 smem[];
  const u32 Llog = LLOG ? (u32)LLOG : p.Llog, Tlog = LLOG ? r16_tlog_c(LLOG) : p.Tlog;
  const u32 L = 1u << Llog, T = 1u << Tlog, TP = T + 1, nthreads = LLOG ? (L * T) / EPT : blockDim.x;   // nthreads * EPT == L * T
  const u32 num_radix = LLOG ? r16_radix_count_c(LLOG) : p.num_radix, radix_logs = LLOG ? r16_radix_packed_c(LLOG) : p.radix_logs;
  u64* S = smem; u64* TW = smem + (size_t)L * TP;   // TW[i] = w_L^(+-i), full circle
  const u32 tid = threadIdx.x, tile = blockIdx.x;
  // transform tr = (group, sub): group selects the source polynomial, sub the coset.  coset_map (4 bits per entry, 0 = identity)
  // lets a job compute a subset of the cosets: table / output slot of sub is map[sub], the output has dst_cosets slots per group.
  // With grp_fast the launch order walks the groups (columns) of one coset first, so that the coset's direct twiddle table is
  // re-read from the L2 and not from HBM.
  const u32 ngrp = gridDim.y / p.src_div;
  const u32 grp = p.grp_fast ? blockIdx.y % ngrp : blockIdx.y / p.src_div, sub = p.grp_fast ? blockIdx.y / ngrp : blockIdx.y % p.src_div;
  const u32 tr = grp * p.src_div + sub;
  const u32 coset = p.coset_map ? (u32)((p.coset_map >> (4 * sub)) & 15) : sub;
  u64* dst = p.dst + (p.coset_map ? (size_t)grp * p.dst_cosets + coset : (size_t)tr) * p.dst_tstride;
  const u64* src = p.src_is_dst ? dst : p.src + (size_t)grp * p.src_tstride;
  // first column of the tile: plain tile * T, or (three-pass plans) split into a slow and a fast part with their own strides
  const u32 thi = p.tile_lo_log ? tile >> p.tile_lo_log : 0u, tlo = p.tile_lo_log ? tile & ((1u << p.tile_lo_log) - 1) : tile;
  const u64 col0 = (u64)thi * p.in_hi_stride + ((u64)tlo << Tlog), col0_out = (u64)thi * p.out_hi_stride + ((u64)tlo << Tlog);

  // ---- load: CH global loads of a thread are issued before their first use (the plain loop over shared-memory stores would
  // serialise them: the compiler cannot prove that `src` does not alias shared memory) ----
  constexpr int CH = 8;
  {
#pragma unroll 4
    for (u32 j = tid; j < L; j += nthreads) { const u32 h = j & (L / 2 - 1); const u64 t = __ldg(p.tw + ((size_t)h << (NTT_TW_LOG - Llog))); TW[j] = j < L / 2 ? t : gl_neg(t); }
    const u32 c = tid & (T - 1), r0 = tid >> Tlog, rstep = nthreads >> Tlog;
    const u64* sp = src + (u64)r0 * p.in_row_stride + col0 + c; const u64 step = (u64)rstep * p.in_row_stride;
    if (p.pre_lo && !p.pre_row) {     // two-level power lookups (no direct tables for this length): rolled, one element at a time
      PowTable pre; pre.lo = p.pre_lo + (size_t)coset * POW_LO; pre.hi = p.pre_hi + (size_t)coset * p.pre_hi_stride;
      const u64 g0 = (u64)r0 * p.in_row_stride + col0 + c;
#pragma unroll 1
      for (int i = 0; i < EPT; i++) S[(r0 + i * rstep) * TP + c] = w_mul(sp[i * step], w_pow_lookup(pre, g0 + i * step));
    } else {
      const u64* pr = p.pre_row ? p.pre_row + ((size_t)coset << Llog) + r0 : nullptr;   // base^(r * in_row_stride); base^(col0 + c) is folded into it_tab
#pragma unroll
      for (int h = 0; h < EPT; h += CH) {
        u64 v[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) v[i] = sp[(h + i) * step];
        if (p.canon_flag) {
          bool bad = false;
#pragma unroll
          for (int i = 0; i < CH; i++) bad |= v[i] >= GL_P;
          if (bad) atomicOr(p.canon_flag, p.canon_bit);
        }
        if (pr) {
#pragma unroll
          for (int i = 0; i < CH; i++) v[i] = w_mul(v[i], __ldg(pr + (h + i) * rstep));      // weak product: the butterflies accept any u64 residue
        }
#pragma unroll
        for (int i = 0; i < CH; i++) S[(r0 + (h + i) * rstep) * TP + c] = v[i];
      }
    }
  }
  __syncthreads();
  u32 ns_log = 0;
  if (LLOG == 10 && XFG_R16_POW2TW) {
    // 1024 = 4 x 16 x 16: the twiddles between the first two passes are 64th roots of unity = powers of two (shifts); only the last pass
    // multiplies by table twiddles: 0.94 full multiplications per element instead of 1.69 with the 16 x 16 x 4 plan
    stockham_pass<2, INV, EPT>(S, TW, 10, Tlog, TP, 0, tid, nthreads);
    stockham_pass<4, INV, EPT, XFG_R16_POW2TW == 1>(S, TW, 10, Tlog, TP, 2, tid, nthreads);
    stockham_pass<4, INV, EPT>(S, TW, 10, Tlog, TP, 6, tid, nthreads);
  } else
#pragma unroll
  for (u32 ps = 0; ps < num_radix; ps++) {
    const u32 lr = (radix_logs >> (4 * ps)) & 15;
    if (EPT >= 32 && lr == 5) stockham_pass<(EPT >= 32 ? 5 : 4), INV, EPT>(S, TW, Llog, Tlog, TP, ns_log, tid, nthreads);
    else if (lr == 4) stockham_pass<4, INV, EPT>(S, TW, Llog, Tlog, TP, ns_log, tid, nthreads);
    else if (lr == 3) stockham_pass<3, INV, EPT>(S, TW, Llog, Tlog, TP, ns_log, tid, nthreads);
    else stockham_pass<2, INV, EPT>(S, TW, Llog, Tlog, TP, ns_log, tid, nthreads);
    ns_log += lr;
  }
  // ---- store ----
  if (p.store_transposed) {
    // Y[(col0 + c) * L + k] = S[k][c] * w_n^((col0 + c) * k)
    // three-pass plans store column lo + 2^swap_lo_log * hi at row hi + 2^swap_hi_log * lo (so that the two later passes run in place)
    auto out_col = [&](u64 col) -> u64 { return p.swap_lo_log ? (col >> p.swap_lo_log) | ((col & ((u64(1) << p.swap_lo_log) - 1)) << p.swap_hi_log) : col; };
    if (p.it_tab) {
      const u64* itab = p.it_tab + (size_t)coset * p.it_tstride;
#pragma unroll
      for (int h = 0; h < EPT; h += CH) {
        u64 w[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) { const u32 e = tid + (h + i) * nthreads, k = e & (L - 1), c = e >> Llog; w[i] = __ldg(itab + (col0 + c) * (u64)L + k); }
#pragma unroll
        for (int i = 0; i < CH; i++) {
          const u32 e = tid + (h + i) * nthreads, k = e & (L - 1), c = e >> Llog;
          dst[out_col(col0 + c) * (u64)L + k] = w_canon(w_mul(S[k * TP + c], w[i]));
        }
      }
    } else {
      PowTable it; it.lo = p.it_lo; it.hi = p.it_hi;
#pragma unroll 1
      for (int i = 0; i < EPT; i++) {
        const u32 e = tid + i * nthreads, k = e & (L - 1), c = e >> Llog;
        u64 v = S[k * TP + c];
        const u64 ex = (col0 + c) * (u64)k;
        v = w_canon(ex ? w_mul(v, w_pow_lookup(it, ex)) : v);
        dst[out_col(col0 + c) * (u64)L + k] = v;
      }
    }
  } else {
    const u32 c = tid & (T - 1), k0 = tid >> Tlog, kstep = nthreads >> Tlog;
    const u64 g0 = (u64)k0 * p.out_row_stride + col0_out + c, gstep = (u64)kstep * p.out_row_stride;
    if (p.post_lo || p.scale != 1) {   // two-level power lookups / separate scale (no direct tables for this length): rolled
      PowTable post; post.lo = p.post_lo ? p.post_lo + (size_t)(tr % p.post_div) * POW_LO : nullptr;
      post.hi = p.post_hi ? p.post_hi + (size_t)(tr % p.post_div) * p.post_hi_stride : nullptr;
#pragma unroll 1
      for (int i = 0; i < EPT; i++) {
        u64 x = S[(k0 + i * kstep) * TP + c];
        const u64 go = g0 + i * gstep;
        if (p.scale != 1) x = gl_mul(x, p.scale); else x = w_canon(x);
        if (post.lo) x = gl_mul(x, pow_lookup(post, go));
        if (p.peer_log) p.peer[go >> p.peer_log][(size_t)tr * (u64(1) << p.peer_log) + (go & ((u64(1) << p.peer_log) - 1))] = x;
        else dst[go] = x;
      }
    } else {
      const u64* pt = p.post_tab ? p.post_tab + (size_t)(tr % p.post_div) * p.post_tstride + g0 : nullptr;
      const u64* rt = p.row_tw;         // three-pass plans: row k of this tile is multiplied by row_tw[hi * k] (one value per row, shared by the T columns)
#pragma unroll
      for (int h = 0; h < EPT; h += CH) {
        u64 w[CH];
        if (rt) {
#pragma unroll
          for (int i = 0; i < CH; i++) w[i] = __ldg(rt + (size_t)thi * (k0 + (h + i) * kstep));
        } else if (pt) {
#pragma unroll
          for (int i = 0; i < CH; i++) w[i] = __ldg(pt + (h + i) * gstep);
        }
#pragma unroll
        for (int i = 0; i < CH; i++) {
          u64 x = S[(k0 + (h + i) * kstep) * TP + c];
          const u64 go = g0 + (h + i) * gstep;
          x = w_canon((pt || rt) ? w_mul(x, w[i]) : x);
          if (p.peer_log) p.peer[go >> p.peer_log][(size_t)tr * (u64(1) << p.peer_log) + (go & ((u64(1) << p.peer_log) - 1))] = x;   // fused all-to-all: store to the row owner
          else dst[go] = x;
        }
      }
    }
  }
}

static size_t r16_smem(u32 Llog, u32 Tlog) { const size_t L = size_t(1) << Llog, TP = (size_t(1) << Tlog) + 1; return (L * TP + L) * sizeof(u64); }
// 4 columns per 2^10-point tile: measured 1.36 ms for the 2^20 trace LDE, against 1.84 ms with 2 columns (half-used sectors) and 1.33 ms with 8 (2 CTAs/SM)
static u32 r16_tlog(u32 Llog) { return r16_tlog_c(Llog); }                       // 4096 elements per CTA (more with 2^11, 2^12-point tiles): 4+ CTAs per SM
// Elements per thread: 16.  A 32-element variant (radix-32 passes, 1024 = 32 x 32, one pass fewer) was measured and is slower
// (136 registers -> 12 warps/SM: LDE of the trace 2.41 ms vs 1.76 ms), so only EPT = 16 is instantiated.
static int r16_ept() { return 16; }
static void r16_radices(u32 Llog, u32 ept, u32& count, u32& packed) {
  // 16 elements per thread: 8:(4,4) 9:(4,3,2) 10:(4,4,2) 11:(4,4,3) 12:(4,4,4);  32 per thread: 8:(5,3) 9:(5,4) 10:(5,5) 11:(4,4,3) 12:(4,4,4)
  static const u32 t16[5][3] = {{4, 4, 0}, {4, 3, 2}, {4, 4, 2}, {4, 4, 3}, {4, 4, 4}};
  static const u32 t32[5][3] = {{5, 3, 0}, {5, 4, 0}, {5, 5, 0}, {4, 4, 3}, {4, 4, 4}};
  const u32* r = (ept == 32 ? t32 : t16)[Llog - 8]; count = r[2] ? 3 : 2; packed = r[0] | (r[1] << 4) | (r[2] << 8);
}
template <bool INV> static void launch_r16_t(cudaStream_t st, const NttPass& p, dim3 grid, u32 threads, size_t sm, bool standard) {
  if (standard) {
    switch (p.Llog) {
      case 8: ntt_pass_r16<INV, 16, 512, 8><<<grid, threads, sm, st>>>(p); return;
      case 9: ntt_pass_r16<INV, 16, 512, 9><<<grid, threads, sm, st>>>(p); return;
      case 10: ntt_pass_r16<INV, 16, 512, 10><<<grid, threads, sm, st>>>(p); return;
      case 11: ntt_pass_r16<INV, 16, 512, 11><<<grid, threads, sm, st>>>(p); return;
      case 12: ntt_pass_r16<INV, 16, 1024, 12><<<grid, threads, sm, st>>>(p); return;
      default: break;
    }
  }
  if (threads > 512) ntt_pass_r16<INV, 16, 1024, 0><<<grid, threads, sm, st>>>(p);
  else ntt_pass_r16<INV, 16, 512, 0><<<grid, threads, sm, st>>>(p);
}
static void launch_r16(cudaStream_t st, NttPass p, bool inverse, u32 tiles, u32 batch) {
  const u32 ept = r16_ept();
  r16_radices(p.Llog, ept, p.num_radix, p.radix_logs);
  const u32 threads = (1u << (p.Llog + p.Tlog)) / ept; const size_t sm = r16_smem(p.Llog, p.Tlog);
  // 2^12-point tiles need 1024 threads (64 registers); everything else runs 256-512 threads
  const bool standard = ept == 16 && p.Tlog == r16_tlog_c(p.Llog) && p.radix_logs == r16_radix_packed_c(p.Llog) && p.num_radix == r16_radix_count_c(p.Llog);
  if (inverse) launch_r16_t<true>(st, p, dim3(tiles, batch), threads, sm, standard); else launch_r16_t<false>(st, p, dim3(tiles, batch), threads, sm, standard);
  XFG_LAUNCHED(1);
}

size_t ntt_pass_smem(u32 Llog, u32 Tlog) {
  size_t L = size_t(1) << Llog, T = size_t(1) << Tlog, TP = T > 1 ? T + 1 : 1;
  return (L * TP + L / 2) * sizeof(u64);
}

// Opt-in dynamic shared memory sizes are per device: called by every xfg_create for its own device (and lazily by ntt_batch).
void ntt_init(bool force) {
  static thread_local bool done = false;
  if (done && !force) return;
  cudaFuncSetAttribute(ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_pass_smem(12, 2));
  cudaFuncSetAttribute(ntt_pass_r16<false, 16, 1024, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r16_smem(12, 2));
  cudaFuncSetAttribute(ntt_pass_r16<true, 16, 1024, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r16_smem(12, 2));
  cudaFuncSetAttribute(ntt_pass_r16<false, 16, 512, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r16_smem(11, 2));
  cudaFuncSetAttribute(ntt_pass_r16<true, 16, 512, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r16_smem(11, 2));
  cudaFuncSetAttribute(ntt_pass_r16<false, 16, 1024, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r16_smem(12, 2));
  cudaFuncSetAttribute(ntt_pass_r16<true, 16, 1024, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r16_smem(12, 2));
  cudaFuncSetAttribute(ntt_pass_r16<false, 16, 512, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r16_smem(11, 2));
  cudaFuncSetAttribute(ntt_pass_r16<true, 16, 512, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r16_smem(11, 2));
  cudaFuncSetAttribute(ntt_pass_r16<false, 16, 512, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r16_smem(10, r16_tlog_c(10)));
  cudaFuncSetAttribute(ntt_pass_r16<true, 16, 512, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)r16_smem(10, r16_tlog_c(10)));
  done = true;
}

// Enqueue a batch of `batch` length-2^ln transforms.  src != dst is required when ln > NTT_SINGLE_MAX_LOG.
void ntt_batch(cudaStream_t st, const NttTables& tb, const NttJob& job) {
  ntt_init(false);
  const u32 ln = job.ln;
  NttPass p{};
  p.tw = job.inverse ? tb.tw_inv : tb.tw_fwd;
  p.src_div = job.src_div ? job.src_div : 1;
  p.src_tstride = job.src_tstride; p.dst_tstride = job.dst_tstride;
  p.pre_lo = job.pre_lo; p.pre_hi = job.pre_hi; p.pre_hi_stride = job.pre_hi_stride;
  p.post_div = job.post_div ? job.post_div : 1;
  p.coset_map = job.coset_map; p.dst_cosets = job.dst_cosets;
  p.canon_flag = job.canon_flag; p.canon_bit = job.canon_bit;
  if (ln <= NTT_SINGLE_MAX_LOG) {
    p.src = job.src; p.dst = job.dst; p.Llog = ln; p.Tlog = 0; p.in_row_stride = 1; p.out_row_stride = 1;
    p.store_transposed = 0; p.scale = job.scale;
    p.post_lo = job.post_lo; p.post_hi = job.post_hi; p.post_hi_stride = job.post_hi_stride;
    p.peer_log = job.peer_log; for (int i = 0; i < NTT_MAX_PEERS; i++) p.peer[i] = job.peer[i];
    ntt_pass<<<dim3(1, job.batch), NTT_THREADS, ntt_pass_smem(ln, 0), st>>>(p); XFG_LAUNCHED(1);
    return;
  }
  // four-step: n = n1 * n2 with n2 = 2^l2 (pass A length), n1 = 2^l1 (pass B length);
  // three-pass (from 2^24 points, needs the direct tables): n = n1 n2 n3 with n3 = 2^(ln-16) (pass A), n2 = n1 = 2^8 (two in-place passes)
  // direct twiddle tables (NttTables): picked when the job's lookup tables are the ones the direct tables were built from
  const bool tab_coset = tb.d_ln == ln && !job.inverse && job.scale == 1 && job.pre_lo && job.pre_lo == tb.d_pre_id && !job.post_lo;
  const bool tab_inv = tb.d_ln == ln && job.inverse && !job.pre_lo && job.scale == tb.d_scale && (!job.post_lo || job.post_lo == tb.d_post_id);
  const bool three = ln >= NTT_THREE_PASS_MIN_LOG && (tab_coset || tab_inv) && tb.d_rtw_fwd && tb.d_l2 == ntt_pass_a_log(ln);
  const u32 l2 = three ? ntt_pass_a_log(ln) : ln / 2, l1 = ln - l2;   // jobs the direct tables do not cover keep the four-step plan with the two-level lookups
  const bool fast = l2 >= 8;            // register-radix Stockham tiles (2^8 .. 2^12 points)
  const bool have_direct = fast && tb.d_ln == ln && tb.d_l2 == l2;
  const bool dir_coset = have_direct && tab_coset, dir_inv = have_direct && tab_inv;
  if (three) {
    // j = j1 + 2^8 j2 + 2^16 j3, k = k3 + n3 k2 + 2^8 n3 k1.
    // A : rows j3 (stride 2^16), T consecutive columns c = j1 + 2^8 j2; x w_n^(c k3) (and the coset / 1/n factors); stored at (j2 + 2^8 j1) n3 + k3
    p.src = job.src; p.dst = job.dst; p.Llog = l2; p.in_row_stride = u64(1) << 16; p.out_row_stride = 0; p.store_transposed = 1; p.scale = 1;
    p.swap_lo_log = 8; p.swap_hi_log = 8;
    if (dir_coset) { p.pre_row = tb.d_pre_row; p.it_tab = tb.d_it_coset; p.it_tstride = u64(1) << ln; p.grp_fast = 1; }
    else { p.it_tab = tb.d_it_inv; p.it_tstride = 0; }
    p.Tlog = r16_tlog(l2); launch_r16(st, p, job.inverse, (1u << 16) >> p.Tlog, job.batch);
    // B1: fixed j1 (hi), rows j2 (stride n3), T consecutive k3, in place; x w_65536^(j1 k2)
    p.src = job.dst; p.src_is_dst = 1; p.pre_lo = nullptr; p.pre_hi = nullptr; p.pre_row = nullptr; p.it_tab = nullptr; p.canon_flag = nullptr;
    p.swap_lo_log = 0; p.swap_hi_log = 0; p.store_transposed = 0; p.scale = 1; p.post_lo = nullptr; p.post_hi = nullptr;
    p.Llog = 8; p.Tlog = r16_tlog(8);
    p.tile_lo_log = l2 - p.Tlog; p.in_hi_stride = p.out_hi_stride = u64(1) << (l2 + 8); p.in_row_stride = p.out_row_stride = u64(1) << l2;
    p.row_tw = job.inverse ? tb.d_rtw_inv : tb.d_rtw_fwd;
    launch_r16(st, p, job.inverse, 256u << p.tile_lo_log, job.batch);
    // B2: fixed k2 (hi), rows j1 (stride 2^8 n3), T consecutive k3, in place: output index k3 + n3 k2 + 2^8 n3 k1 (natural order)
    p.row_tw = nullptr; p.in_hi_stride = p.out_hi_stride = u64(1) << l2; p.in_row_stride = p.out_row_stride = u64(1) << (l2 + 8);
    if (dir_inv && job.post_lo) { p.post_tab = tb.d_post; p.post_tstride = u64(1) << ln; }
    p.peer_log = job.peer_log; for (int i = 0; i < NTT_MAX_PEERS; i++) p.peer[i] = job.peer[i];
    launch_r16(st, p, job.inverse, 256u << p.tile_lo_log, job.batch);
    return;
  }
  const u32 Tlog = l1 > 11 ? 2 : 3;     // 2^12-point tiles only fit 4 columns
  // pass A
  p.src = job.src; p.dst = job.dst; p.Llog = l2; p.Tlog = Tlog; p.in_row_stride = u64(1) << l1; p.out_row_stride = 0;
  p.store_transposed = 1; p.scale = 1;
  p.it_lo = job.inverse ? tb.wn_inv.lo : tb.wn_fwd.lo; p.it_hi = job.inverse ? tb.wn_inv.hi : tb.wn_fwd.hi;
  if (dir_coset) { p.pre_row = tb.d_pre_row; p.it_tab = tb.d_it_coset; p.it_tstride = u64(1) << ln; p.grp_fast = 1; }
  if (dir_inv) { p.it_tab = tb.d_it_inv; p.it_tstride = 0; }
  if (fast) { p.Tlog = r16_tlog(l2); launch_r16(st, p, job.inverse, (1u << l1) >> p.Tlog, job.batch); }
  else { ntt_pass<<<dim3((1u << l1) >> Tlog, job.batch), NTT_THREADS, ntt_pass_smem(l2, Tlog), st>>>(p); XFG_LAUNCHED(1); }
  // pass B (in place on dst)
  p.src = job.dst; p.dst = job.dst; p.src_is_dst = 1;                      // pass B: in place on the slot pass A wrote
  p.pre_lo = nullptr; p.pre_hi = nullptr; p.pre_row = nullptr; p.it_tab = nullptr; p.canon_flag = nullptr;
  p.Llog = l1; p.in_row_stride = u64(1) << l2; p.out_row_stride = u64(1) << l2; p.store_transposed = 0; p.scale = job.scale;
  p.post_lo = job.post_lo; p.post_hi = job.post_hi; p.post_hi_stride = job.post_hi_stride;
  if (dir_inv) {   // 1/n already applied by d_it_inv
    p.scale = 1;
    if (job.post_lo) { p.post_tab = tb.d_post; p.post_tstride = u64(1) << ln; p.post_lo = nullptr; p.post_hi = nullptr; }
  }
  p.peer_log = job.peer_log; for (int i = 0; i < NTT_MAX_PEERS; i++) p.peer[i] = job.peer[i];
  if (fast) { p.Tlog = r16_tlog(l1); launch_r16(st, p, job.inverse, (1u << l2) >> p.Tlog, job.batch); }
  else { ntt_pass<<<dim3((1u << l2) >> Tlog, job.batch), NTT_THREADS, ntt_pass_smem(l1, Tlog), st>>>(p); XFG_LAUNCHED(1); }
}

// ---- direct twiddle tables (NttTables::d_*) ----
__global__ void fill_it_kernel(u64* __restrict__ out, u32 ln, u32 l2, PowTable wn, u64 scale, PowTable base) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >> ln) return;
  const u64 j1 = i >> l2, k2 = i & ((u64(1) << l2) - 1);
  u64 v = gl_mul(pow_lookup(wn, j1 * k2), scale);
  if (base.lo) v = gl_mul(v, pow_lookup(base, j1));
  out[i] = v;
}
__global__ void fill_pow_kernel(u64* __restrict__ out, size_t count, u32 shift, PowTable base) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = pow_lookup(base, (u64)i << shift);
}
size_t ntt_direct_words(u32 ln, u32 cosets, u32 posts) {
  if (ln < NTT_DIRECT_MIN_LOG || ln > NTT_DIRECT_MAX_LOG) return 0;
  const size_t n = size_t(1) << ln, n2 = size_t(1) << ntt_pass_a_log(ln);
  return n + cosets * n + cosets * n2 + posts * n + (ln >= NTT_THREE_PASS_MIN_LOG ? (size_t(2) << 16) : 0);
}
void ntt_build_direct(NttTables& tb, u32 ln, u64* storage, u64 scale, const u64* pre_lo, const u64* pre_hi, u32 pre_hi_stride, u32 cosets,
                      const u64* post_lo, const u64* post_hi, u32 post_hi_stride, u32 posts) {
  if (!ntt_direct_words(ln, cosets, posts)) return;
  const size_t n = size_t(1) << ln; const u32 l2 = ntt_pass_a_log(ln), l1 = ln - l2; const size_t n2 = size_t(1) << l2;
  const unsigned blocks = (unsigned)(n / 256);
  u64* it_inv = storage; u64* it_coset = it_inv + n; u64* pre_row = it_coset + cosets * n; u64* post = pre_row + cosets * n2;
  fill_it_kernel<<<blocks, 256>>>(it_inv, ln, l2, tb.wn_inv, scale, PowTable{nullptr, nullptr});
  for (u32 c = 0; c < cosets; c++) {
    const PowTable base{pre_lo + (size_t)c * POW_LO, pre_hi + (size_t)c * pre_hi_stride};
    fill_it_kernel<<<blocks, 256>>>(it_coset + (size_t)c * n, ln, l2, tb.wn_fwd, 1, base);
    fill_pow_kernel<<<(unsigned)((n2 + 255) / 256), 256>>>(pre_row + (size_t)c * n2, n2, l1, base);
  }
  for (u32 c = 0; c < posts; c++)
    fill_pow_kernel<<<blocks, 256>>>(post + (size_t)c * n, n, 0, PowTable{post_lo + (size_t)c * POW_LO, post_hi + (size_t)c * post_hi_stride});
  tb.d_it_inv = it_inv; tb.d_it_coset = it_coset; tb.d_pre_row = pre_row; tb.d_post = posts ? post : nullptr;
  tb.d_pre_id = pre_lo; tb.d_post_id = posts ? post_lo : nullptr; tb.d_scale = scale; tb.d_ln = ln; tb.d_l2 = l2;
  if (ln >= NTT_THREE_PASS_MIN_LOG) {   // w_65536^(+-e) = w_n^(+-e * n / 65536), e < 2^16: the twiddles between the two in-place passes
    u64* rf = post + (size_t)posts * n; u64* ri = rf + (size_t(1) << 16);
    fill_pow_kernel<<<256, 256>>>(rf, size_t(1) << 16, ln - 16, tb.wn_fwd);
    fill_pow_kernel<<<256, 256>>>(ri, size_t(1) << 16, ln - 16, tb.wn_inv);
    tb.d_rtw_fwd = rf; tb.d_rtw_inv = ri;
  }
}

}  // namespace xfg
