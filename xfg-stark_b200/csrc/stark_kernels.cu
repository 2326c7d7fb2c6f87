// stark_kernels.cu — the elementwise stages of the burn-mint proof for sm_100a: AIR constraint evaluation, composition
// coefficients, out-of-domain evaluation, DEEP composition and FRI folding.  All streaming kernels over the coset-major LDE
// (LDE row i = 8m + k at [k*n + m]), one field element per lane, every global access coalesced along m.
//
// Replaces winter-prover 0.8.3 `DefaultConstraintEvaluator::evaluate` + `ConstraintEvaluationTable::combine`,
// `CompositionPoly::new`, `TracePolyTable::get_ood_frame`, `DeepCompositionPoly::{add_trace_polys, add_composition_poly,
// evaluate}` and winter-fri `FriProver::build_layers` / `folding::apply_drp`, together
// with the reference's `Air::evaluate_transition` / `get_assertions` (src/burn_mint_air.rs:335-395).
#include "stark_kernels.cuh"
#include "launch.cuh"
#include "field_weak.cuh"

namespace xfg {

template <int D> __device__ __forceinline__ Ext<D> ld_ext(const u64 (*p)[2], int i) { return Ext<D>(p[i][0], p[i][1]); }
template <int D> __device__ __forceinline__ Ext<D> ld_ext1(const u64* p) { return Ext<D>(p[0], p[1]); }

// 8-byte asynchronous copy global -> shared (LDGSTS): the prefetches of the constraint and DEEP kernels (no registers held while the load is in flight)
__device__ __forceinline__ void cp_async8(u64* smem_dst, const u64* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((u32)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// in-register Montgomery batch inversion of K non-zero base-field values
template <int K> __device__ __forceinline__ void batch_inv(u64 (&v)[K]) {
  u64 pre[K]; u64 acc = 1;
#pragma unroll
  for (int i = 0; i < K; i++) { pre[i] = acc; acc = gl_mul(acc, v[i]); }
  acc = w_inv(acc);
#pragma unroll
  for (int i = K - 1; i >= 0; i--) { u64 t = gl_mul(pre[i], acc); acc = gl_mul(acc, v[i]); v[i] = t; }
}

// ------------------------------------------------------------------------------------------------------------------
// evaluate_constraints: one thread evaluates CE_PTS points of the constraint-evaluation coset k' (LDE coset 4k').
//   H(x) = T(x) (x - g^(n-1)) / (x^n - 1) + B0(x) / (x - 1) + B1(x) / (x - g^(n-1))      (A.8)
// out: [limb][k'][m], k' < 2
// ------------------------------------------------------------------------------------------------------------------
#ifndef XFG_CE_PTS
#define XFG_CE_PTS 8      // 8 points per thread share one inversion (30 % of the 4-point kernel was the inversion): 0.210 -> 0.188 ms at 2^20
#endif
// XFG_DEEP_CTA_INV / XFG_CE_CTA_INV: the one inversion per thread (72 dependent multiplications, a fifth to a third of these kernels' instructions when
// every warp runs its own chain) becomes one inversion per BLOCK (cta_inv, field_weak.cuh).  Measured at 2^20 rows / quadratic (B200, round 2):
//   DEEP        0.602 -> 0.561 ms (128 threads; 256 threads / 2 blocks per SM: 0.565)                                  -> kept (default 1)
//   constraints 0.175 -> 0.182 ms (64 threads), 0.175 (128 threads), 0.176 (128 threads, 4 points per thread)          -> not kept (default 0):
//   that kernel is bound by latency (ALU pipe 57 % busy), so the instructions saved are paid back by the two barriers
#ifndef XFG_DEEP_CTA_INV
#define XFG_DEEP_CTA_INV 1
#endif
#ifndef XFG_CE_CTA_INV
#define XFG_CE_CTA_INV 0
#endif
#ifndef XFG_CE_THREADS
#define XFG_CE_THREADS 64
#endif
static constexpr int CE_PTS = XFG_CE_PTS, CE_THREADS = XFG_CE_THREADS;
#ifndef XFG_CE_MINB
#define XFG_CE_MINB 8
#endif
#ifndef XFG_CE_PREFETCH
#define XFG_CE_PREFETCH 1
#endif
// The loops over the CE_PTS points are NOT unrolled and the per-point intermediates live in shared memory ([point][word][thread],
// conflict-free): the fully unrolled version was 160 KB of SASS and stalled on instruction fetch (ncu: no_instruction).
template <int D>
__global__ void __launch_bounds__(CE_THREADS, XFG_CE_MINB) constraint_kernel(const u64* __restrict__ lde, u32 ln, const AirParams* __restrict__ airp, const ProofState* __restrict__ ps,
                                                                 PowTable wn, u64 s_k0, u64 s_k1, u64 zinv0, u64 zinv1, u64* __restrict__ out, size_t out_tstride) {
  extern __shared__ u64 ce_dsm[];
  u64 (*sh)[2 * D + 2][CE_THREADS] = reinterpret_cast<u64 (*)[2 * D + 2][CE_THREADS]>(ce_dsm);   // [CE_PTS points][u (D), w (D), d, prefix][thread]
  __shared__ u64 sc[(XFG_NUM_TRANSITION + XFG_NUM_ASSERTIONS) * 2];
  const AirParams air = *airp;                            // AIR constants live in device memory so that the launch is CUDA-graph replayable
  const size_t n = size_t(1) << ln, N = 8 * n;
  const u32 kp = blockIdx.y, k = kp * 4, tid = threadIdx.x;
  const size_t per = n / CE_PTS, t = (size_t)blockIdx.x * blockDim.x + tid;
  if (tid < (XFG_NUM_TRANSITION + XFG_NUM_ASSERTIONS) * 2) sc[tid] = (&ps->tcoef[0][0])[tid];   // tcoef[7][2] then bcoef[8][2], contiguous
  __syncthreads();
  const bool active = t < per;
  if (!XFG_CE_CTA_INV && !active) return;
  const u64 sk = kp ? s_k1 : s_k0, zinv = kp ? zinv1 : zinv0;
  const u64 large_burn = gl_mul(XFG_STD_BURN, 1000);
  u64 acc = 1;
  // XFG_CE_PREFETCH: the 8 loads of point q + 1 run as cp.async (global -> shared) under the arithmetic of point q (ncu source page: 28 % of the plain
  // kernel's warp samples were `long_scoreboard` on the first use of a loaded value)
#if XFG_CE_PREFETCH
  __shared__ u64 pf[XFG_TRACE_WIDTH + 1][CE_THREADS];
  auto fetch = [&](int q) {
    const size_t m = t + q * per, mn = (m + 1) & (n - 1);
#pragma unroll
    for (int j = 0; j < XFG_TRACE_WIDTH; j++) cp_async8(&pf[j][tid], lde + (size_t)j * N + (size_t)k * n + m);
    cp_async8(&pf[XFG_TRACE_WIDTH][tid], lde + (size_t)4 * N + (size_t)k * n + mn);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (active) fetch(0);
#endif
#pragma unroll 1
  for (int q = 0; q < (active ? CE_PTS : 0); q++) {
    const size_t m = t + q * per, mn = (m + 1) & (n - 1);
    u64 c[XFG_TRACE_WIDTH];
#if XFG_CE_PREFETCH
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int j = 0; j < XFG_TRACE_WIDTH; j++) c[j] = pf[j][tid];
    const u64 nxt4 = pf[XFG_TRACE_WIDTH][tid];
    if (q + 1 < CE_PTS) fetch(q + 1);
    (void)mn;
#else
#pragma unroll
    for (int j = 0; j < XFG_TRACE_WIDTH; j++) c[j] = lde[(size_t)j * N + (size_t)k * n + m];
    const u64 nxt4 = lde[(size_t)4 * N + (size_t)k * n + mn];
#endif
    // src/burn_mint_air.rs:356-377
    u64 r[XFG_NUM_TRANSITION];
    r[0] = gl_mul(gl_sub(c[0], XFG_STD_BURN), gl_sub(c[0], large_burn));
    r[1] = gl_sub(c[1], c[0]);
    r[2] = gl_sub(c[2], air.txn);
    r[3] = gl_sub(c[3], air.rcpt);
    const u64 d = gl_sub(nxt4, c[4]); r[4] = gl_mul(d, gl_sub(d, 1));
    r[5] = gl_sub(c[5], air.nullifier);
    r[6] = gl_sub(c[6], air.commitment);
    // random linear combinations as un-reduced dot products (DotAcc), one reduction per limb
    DotAcc ta[D], ba[D];
#pragma unroll
    for (int j = 0; j < XFG_NUM_TRANSITION; j++) for (int l = 0; l < D; l++) ta[l].fma(sc[2 * j + l], r[j]);
    // src/burn_mint_air.rs:383-394 in Winterfell's sorted order: step-0 columns 0..6, then (column 4, step n-1)
#pragma unroll
    for (int j = 0; j < XFG_TRACE_WIDTH; j++) { const u64 dj = gl_sub(c[j], air.assert0[j]); for (int l = 0; l < D; l++) ba[l].fma(sc[2 * (XFG_NUM_TRANSITION + j) + l], dj); }
    Ext<D> ts, bs;
#pragma unroll
    for (int l = 0; l < D; l++) { ts.set_limb(l, ta[l].result()); bs.set_limb(l, ba[l].result()); }
    const Ext<D> b1 = mul_base(Ext<D>(sc[2 * (XFG_NUM_TRANSITION + XFG_TRACE_WIDTH)], sc[2 * (XFG_NUM_TRANSITION + XFG_TRACE_WIDTH) + 1]), gl_sub(c[4], XFG_FINAL_STATE));
    const u64 x = gl_mul(sk, pow_lookup(wn, m));
    const u64 xm1 = gl_sub(x, 1), xml = gl_sub(x, air.g_last);
    //   H = T (x - g_last) / (x^n - 1)  +  [ B0 (x - g_last) + B1 (x - 1) ] / [ (x - 1)(x - g_last) ]
    const Ext<D> u = mul_base(ts, gl_mul(xml, zinv)), w = mul_base(bs, xml) + mul_base(b1, xm1);
    const u64 dd = gl_mul(xm1, xml);
#pragma unroll
    for (int l = 0; l < D; l++) { sh[q][l][tid] = u.limb(l); sh[q][D + l][tid] = w.limb(l); }
    sh[q][2 * D][tid] = dd; sh[q][2 * D + 1][tid] = acc;
    acc = gl_mul(acc, dd);
  }
#if XFG_CE_CTA_INV
  __shared__ u64 invbuf[CE_THREADS];
  acc = cta_inv<CE_THREADS>(acc, invbuf);      // a thread without work hands in 1
  if (!active) return;
#else
  acc = w_inv(acc);
#endif
#pragma unroll 1
  for (int q = CE_PTS - 1; q >= 0; q--) {
    const size_t m = t + q * per;
    const u64 dinv = gl_mul(sh[q][2 * D + 1][tid], acc); acc = gl_mul(acc, sh[q][2 * D][tid]);
#pragma unroll
    for (int l = 0; l < D; l++) out[(size_t)(l * 2 + kp) * out_tstride + m] = gl_add(sh[q][l][tid], gl_mul(sh[q][D + l][tid], dinv));
  }
}

static size_t ce_smem(int D) { return (size_t)CE_PTS * (2 * D + 2) * CE_THREADS * sizeof(u64); }
void launch_constraints(cudaStream_t st, int D, const u64* lde, u32 ln, const AirParams* air, const ProofState* ps, PowTable wn,
                        u64 s_k0, u64 s_k1, u64 zinv0, u64 zinv1, u64* out, size_t out_tstride) {
  const size_t per = (size_t(1) << ln) / CE_PTS; dim3 grid((unsigned)((per + CE_THREADS - 1) / CE_THREADS), 2);
  if (D == 1) constraint_kernel<1><<<grid, CE_THREADS, ce_smem(1), st>>>(lde, ln, air, ps, wn, s_k0, s_k1, zinv0, zinv1, out, out_tstride);
  else constraint_kernel<2><<<grid, CE_THREADS, ce_smem(2), st>>>(lde, ln, air, ps, wn, s_k0, s_k1, zinv0, zinv1, out, out_tstride);
  XFG_LAUNCHED(1);
}

// ------------------------------------------------------------------------------------------------------------------
// composition coefficients: after the two size-n inverse NTTs (un-scaled by the coset offsets), the size-2n interpolant is
//   h[j] = (A0[j] + A1[j]) / 2  for j < n   and   h[n + j] = (A0[j] - A1[j]) / (2 * 7^n), which must vanish (degree check).
// a: [limb][2][n] -> h: [limb][n]
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) combine_kernel(const u64* __restrict__ a, u32 ln, int D, u64 inv2, u64* __restrict__ h, ProofState* ps) {
  const size_t n = size_t(1) << ln, j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  bool bad = false;
  for (int l = 0; l < D; l++) {
    u64 a0 = a[(size_t)l * 2 * n + j], a1 = a[(size_t)l * 2 * n + n + j];
    h[(size_t)l * n + j] = gl_mul(gl_add(a0, a1), inv2);
    bad |= (a0 != a1);
  }
  if (bad) atomicOr(&ps->error_flags, ERR_FLAG_DEGREE);
}
void launch_combine(cudaStream_t st, const u64* a, u32 ln, int D, u64 inv2, u64* h, ProofState* ps) {
  const size_t n = size_t(1) << ln;
  combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, ln, D, inv2, h, ps); XFG_LAUNCHED(1);
}

// ------------------------------------------------------------------------------------------------------------------
// out-of-domain evaluation: poly p (base-field coefficients) at z and z*g.  Thread t sums c[t + i*TOT] z^(t + i*TOT) by a
// Horner pass in z^TOT; block partials are added by the transcript kernel.  partial: [poly][block][point][limb]
// ------------------------------------------------------------------------------------------------------------------
#ifndef XFG_OOD_MLP
#define XFG_OOD_MLP 4
#endif
#ifndef XFG_OOD_SHARE
#define XFG_OOD_SHARE 0
#endif
#ifndef XFG_OOD_LOCAL_POW
#define XFG_OOD_LOCAL_POW 1
#endif
static constexpr u32 OOD_THREADS = 256;
template <int D>
__global__ void __launch_bounds__(256) ood_kernel(const u64* __restrict__ trace_coef, const u64* __restrict__ h_coef, u32 ln, u32 width, u32 polys_per_block,
                                                   const ProofState* __restrict__ ps, u64* __restrict__ partial) {
  const size_t n = size_t(1) << ln;
  const u32 nb = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t TOT = (size_t)nb * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + tid;
  const u32 P = width + D, p0 = blockIdx.y * polys_per_block, p1 = min(P, p0 + polys_per_block);
  // powers of the two evaluation points shared by the block: sq[w][b] = pt_w^(2^b), b < 32 (one lane per (w, b) chain would be
  // serial anyway: thread w squares 31 times), then every thread assembles pt^t and pt^TOT from the set bits of t / TOT
  __shared__ u64 sq[2][32][2];
  const int nbits = (int)ln < 1 ? 1 : (int)ln;      // every exponent is < n = 2^ln
  if (tid < 2) {
    Ext<D> x = ld_ext1<D>(tid == 0 ? ps->z : ps->zg);
    for (int b = 0; b < nbits; b++) { sq[tid][b][0] = x.limb(0); sq[tid][b][1] = D == 2 ? x.limb(1) : 0; x = x * x; }
  }
  __syncthreads();
  auto pw = [&](int w, size_t e) { Ext<D> r(1); for (int b = 0; b < nbits; b++) if ((e >> b) & 1) r = r * Ext<D>(sq[w][b][0], sq[w][b][1]); return r; };
  // table of (pt^TOT)^i, i < n/TOT, so that a thread's share  sum_i c[t + i TOT] pt^(i TOT)  is a plain dot product of base-field
  // coefficients with table limbs: accumulated un-reduced (DotAcc), 4 cheap fma per coefficient instead of 2 extension Horner steps
  __shared__ u64 tab[OOD_MAX_STEPS][2][2];
  const size_t steps = n > TOT ? n / TOT : 1;
  for (size_t i = tid; i < steps; i += blockDim.x) for (int w = 0; w < 2; w++) { const Ext<D> v = pw(w, i * TOT); tab[i][w][0] = v.limb(0); tab[i][w][1] = D == 2 ? v.limb(1) : 0; }
#if XFG_OOD_LOCAL_POW
  __shared__ u64 lp[2][OOD_THREADS][2];      // pt_w^i, i < 256
  __shared__ u64 bp[2][2];                   // pt_w^(first exponent of this block)
  if (tid >= OOD_THREADS - 2) {              // the last two threads: they have no table entry to compute unless steps = 256
    const u32 w = tid - (OOD_THREADS - 2);
    const Ext<D> b = pw((int)w, (size_t)blockIdx.x * OOD_THREADS);
    bp[w][0] = b.limb(0); bp[w][1] = D == 2 ? b.limb(1) : 0;
    lp[w][0][0] = 1; lp[w][0][1] = 0;
  }
#endif
  __syncthreads();
  // the powers pt^t of this thread serve every polynomial of the block's group (blockIdx.y): wide traces amortise the set-up above
  Ext<D> pt[2];
#if XFG_OOD_LOCAL_POW
  // pt^t = pt^(block base) * pt^tid.  The 256 powers pt^tid are built by doubling - thread i multiplies ONCE, at step floor(log2 i):
  // lp[i] = lp[i - 2^s] * pt^(2^s) - and the base by one thread per point: 2 extension multiplications per thread and point instead of one per set
  // bit of t (a warp ran ~14 per point, a third of the kernel's instructions).  Exponents beyond n are never used (t < n is checked below).
#pragma unroll 1
  for (int sft = 0; (1 << sft) < OOD_THREADS; sft++) {
    if ((tid >> sft) == 1 && sft < nbits) {
#pragma unroll
      for (int w = 0; w < 2; w++) {
        const Ext<D> v = Ext<D>(lp[w][tid - (1u << sft)][0], lp[w][tid - (1u << sft)][1]) * Ext<D>(sq[w][sft][0], sq[w][sft][1]);
        lp[w][tid][0] = v.limb(0); lp[w][tid][1] = D == 2 ? v.limb(1) : 0;
      }
    }
    __syncthreads();
  }
  if (t < n) {
#pragma unroll
    for (int w = 0; w < 2; w++) pt[w] = Ext<D>(lp[w][tid][0], lp[w][tid][1]) * Ext<D>(bp[w][0], bp[w][1]);
  }
#else
  if (t < n) { pt[0] = pw(0, t); pt[1] = pw(1, t); }
#endif
  __shared__ u64 red[2][8][4];
#pragma unroll 1
  for (u32 poly = p0; poly < p1; poly++) {
    const u64* c = poly < width ? trace_coef + (size_t)poly * n : h_coef + (size_t)(poly - width) * n;
    u64 r[4] = {0, 0, 0, 0};
    if (t < n) {
      DotAcc d[2][D];
      // XFG_OOD_MLP coefficient loads are issued before their products: the loop is bound by the latency of its dependent global loads, not by
      // bandwidth (75 MB in 0.097 ms) - (8 in flight were tried in round 1: 78 registers, 0.123 ms)
      size_t i = 0;
      for (; i + XFG_OOD_MLP <= steps; i += XFG_OOD_MLP) {
        u64 cv[XFG_OOD_MLP];
#pragma unroll
        for (int q = 0; q < XFG_OOD_MLP; q++) cv[q] = c[t + (i + q) * TOT];
#pragma unroll
        for (int q = 0; q < XFG_OOD_MLP; q++)
#pragma unroll
          for (int w = 0; w < 2; w++) for (int l = 0; l < D; l++) d[w][l].fma(cv[q], tab[i + q][w][l]);
      }
      for (; i < steps; i++) {
        const u64 cv = c[t + i * TOT];
#pragma unroll
        for (int w = 0; w < 2; w++) for (int l = 0; l < D; l++) d[w][l].fma(cv, tab[i][w][l]);
      }
#pragma unroll
      for (int w = 0; w < 2; w++) { Ext<D> a; for (int l = 0; l < D; l++) a.set_limb(l, d[w][l].result()); a = a * pt[w]; for (int l = 0; l < D; l++) r[2 * w + l] = a.limb(l); }
    }
    // block sum: shuffles inside the warps, then 8 warp sums per value (exact arithmetic: any order gives the same element)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (D == 1 && (q & 1)) continue;
      for (int o = 16; o > 0; o >>= 1) r[q] = gl_add(r[q], __shfl_xor_sync(0xFFFFFFFFu, r[q], o));
    }
    u64 (*rb)[4] = red[poly & 1];            // double-buffered: one barrier per polynomial
    if (lane == 0) for (int q = 0; q < 4; q++) rb[warp][q] = r[q];
    __syncthreads();
    if (tid < 4) {
      u64 s = 0;
      for (u32 wv = 0; wv < blockDim.x / 32; wv++) s = gl_add(s, rb[wv][tid]);
      partial[(((size_t)poly * nb + blockIdx.x) * 2 + (tid >> 1)) * 2 + (tid & 1)] = s;
    }
  }
}
u32 ood_num_blocks(u32 ln) { size_t n = size_t(1) << ln; size_t b = n / 256; if (b < 1) b = 1; if (b > OOD_MAX_BLOCKS) b = OOD_MAX_BLOCKS; return (u32)b; }   // 64 x 256 threads per polynomial: Horner chains of n / 16384 steps
void launch_ood(cudaStream_t st, int D, const u64* trace_coef, const u64* h_coef, u32 ln, u32 width, const ProofState* ps, u64* partial) {
  // one polynomial per block while that already gives >= 4 blocks per SM, else groups of polynomials share a block's power tables
  // (sharing more - 2 polynomials per block at width 7 - was measured slower, 0.097 -> 0.118 ms: the kernel is latency-bound and wants blocks)
  const u32 nb = ood_num_blocks(ln), P = width + D;
  u32 ppb = 1; while ((size_t)nb * ((P + ppb - 1) / ppb) > 148 * 4 && ppb < P) ppb++;
#if XFG_OOD_SHARE
  // A/B switch: every block evaluates ALL polynomials on its slice of exponents, so the per-thread power set-up (pt^t) is paid once instead of once per
  // polynomial.  MEASURED SLOWER at 2^20 rows (round 2): 0.089 ms -> 0.140 / 0.111 / 0.130 ms with 128 / 256 / 512 blocks: the kernel wants its 576 short blocks.
  ppb = P;
#endif
  dim3 grid(nb, (P + ppb - 1) / ppb);
  if (D == 1) ood_kernel<1><<<grid, OOD_THREADS, 0, st>>>(trace_coef, h_coef, ln, width, ppb, ps, partial);
  else ood_kernel<2><<<grid, OOD_THREADS, 0, st>>>(trace_coef, h_coef, ln, width, ppb, ps, partial);
  XFG_LAUNCHED(1);
}

// ------------------------------------------------------------------------------------------------------------------
// DEEP composition, pointwise on the LDE coset (A.9; identical field elements to the coefficient-domain computation):
//   D(x) = [ (S(x) - C1)(x - zg) + (S_T(x) - C2)(x - z) ] / [ (x - z)(x - zg) ]
//   S_T = sum_j gamma_j T_j(x),  S = S_T + delta H(x),  C1 = sum_j gamma_j T_j(z) + delta H(z),  C2 = sum_j gamma_j T_j(zg)
// Thread (k, a) computes the 8 points m = a + j*n/8 - exactly the 8 elements of row 8a + k of the first FRI layer - with one
// batched inversion, writes them coset-major and hashes the row into the layer-0 FRI tree.
// ------------------------------------------------------------------------------------------------------------------
#ifndef XFG_DEEP_THREADS
#define XFG_DEEP_THREADS 128
#endif
static constexpr int DEEP_THREADS = XFG_DEEP_THREADS;
#ifndef XFG_DEEP_UNROLL
#define XFG_DEEP_UNROLL 1   // points of the first loop interleaved per iteration (A/B switch).  Measured at 2^20 rows / quadratic (round 2): 1 -> 0.633 ms, 2 -> 0.676, 4 -> 0.692
                            // (128 registers, no spills in all three: the larger loop body costs instruction-cache hits, as the fully unrolled kernel of round 1 did)
#endif
static constexpr int DEEP_UNROLL = XFG_DEEP_UNROLL;
// 1 / (x - w) for a base-field x and an extension point w = (w0, w1):  (x - w)^-1 = conj / norm with
//   u = x - w0,  conj = (u - w1, w1),  norm = u (u - w1) + 2 w1^2           (degree 2; one multiplication)
//   conj = 1,    norm = x - w0                                              (degree 1)
// and for P = (p0, p1):  P * conj = (p0 (u - w1) - 2 p1 w1,  p1 u + p0 w1)  - four products, accumulated un-reduced.
template <int D> struct DeepPoint { u64 w0, w1, m2w1, k2w1sq; };   // w, -2 w1, 2 w1^2
template <int D> __device__ __forceinline__ u64 deep_norm(const DeepPoint<D>& w, u64 x, u64& u) {
  u = gl_sub(x, w.w0);
  if (D == 1) return u;
  return gl_add(gl_mul(u, gl_sub(u, w.w1)), w.k2w1sq);
}
template <int D> __device__ __forceinline__ Ext<D> deep_mul_conj(const DeepPoint<D>& w, u64 u, Ext<D> p) {
  if (D == 1) return p;
  DotAcc c0, c1;
  c0.fma(p.limb(0), gl_sub(u, w.w1)); c0.fma(p.limb(1), w.m2w1);
  c1.fma(p.limb(1), u); c1.fma(p.limb(0), w.w1);
  return Ext<D>(c0.result(), c1.result());
}
// Loops over the 8 points are rolled and the per-point intermediates (numerator, norm product, prefix product) are staged in
// shared memory [point][word][thread]; the unrolled version was 277 KB of SASS, 136 registers, and stalled on instruction fetch.
//   D(x) = P / (x - z) + Q / (x - zg) = [ (P conj_z) n_zg + (Q conj_zg) n_z ] / (n_z n_zg),   P = S_T + delta H - C1,  Q = S_T - C2
// so a point costs one base-field inversion (batched over the thread's 8 points) and no extension-field multiplication
// besides delta * H.
#ifndef XFG_DEEP_MINB
#define XFG_DEEP_MINB 4   // 122 registers, no spills: measured 0.652 ms; 3 (146 regs) 0.664, unconstrained (152) 0.662, 5 (96 regs, 24 B spilled) 0.740, 6-7: 0.69
#endif
// WC = compile-time trace width (the burn-mint AIR: 7, loops over the columns unrolled) or 0 = run-time width `width_rt` (generic
// AIR front-end, up to XFG_AIR_MAX_WIDTH columns); dcoef = width + 1 DEEP coefficients of 2 limbs.
// XFG_DEEP_PREFETCH (burn-mint instantiation, WC != 0): the WC + D values of point j + 1 are fetched with cp.async (global -> shared, no registers) while
// point j is computed; the plain loop consumes its loads right after issuing them (ncu source page: 10 % of the kernel's warp samples sit on the first use
// of a loaded value, `long_scoreboard`, with 4 warps per scheduler to hide it)
#ifndef XFG_DEEP_PREFETCH
#define XFG_DEEP_PREFETCH 1
#endif
template <int D, int WC>
__global__ void __launch_bounds__(DEEP_THREADS, XFG_DEEP_MINB) deep_kernel(const u64* __restrict__ lde, const u64* __restrict__ hlde, u32 ln, const ProofState* __restrict__ ps,
                                                             const u64* __restrict__ dcoef, u32 width_rt,
                                                             PowTable wn, const u64* __restrict__ s_k, u64 w8, u64* __restrict__ deep, Digest* __restrict__ fri_tree0) {
  constexpr int MAXW = WC ? WC : XFG_AIR_MAX_WIDTH;
  const int W = WC ? WC : (int)width_rt;
  extern __shared__ u64 deep_dsm[];
  u64 (*sh)[D + 2][DEEP_THREADS] = reinterpret_cast<u64 (*)[D + 2][DEEP_THREADS]>(deep_dsm);   // [8 points][numerator -> result (D), n_z n_zg, prefix][thread]
  __shared__ u64 sc[2 * (MAXW + 1) + 8];              // dcoef[W + 1][2], then (at 2 (MAXW + 1)) c1, c2, z, zg
  const size_t n = size_t(1) << ln, N = 8 * n, n8 = n / 8;
  const u32 k = blockIdx.y, tid = threadIdx.x; const size_t a = (size_t)blockIdx.x * blockDim.x + tid;
  for (u32 i = tid; i < 2u * (W + 1); i += DEEP_THREADS) sc[i] = dcoef[i];
  if (tid < 8) { const u64* src = tid < 2 ? ps->deep_c1 : tid < 4 ? ps->deep_c2 : tid < 6 ? ps->z : ps->zg; sc[2 * (MAXW + 1) + tid] = src[tid & 1]; }
  __syncthreads();
  const bool active = a < n8;
  if (!XFG_DEEP_CTA_INV && !active) return;
  const u64* cc = sc + 2 * (MAXW + 1);
  const Ext<D> delta(sc[2 * W], sc[2 * W + 1]), c1(cc[0], cc[1]), c2(cc[2], cc[3]);
  const DeepPoint<D> pz{cc[4], cc[5], gl_neg(gl_dbl(cc[5])), gl_dbl(gl_sqr(cc[5]))}, pzg{cc[6], cc[7], gl_neg(gl_dbl(cc[7])), gl_dbl(gl_sqr(cc[7]))};
  u64 x = gl_mul(s_k[k], pow_lookup(wn, active ? a : 0)), acc = 1;      // x_j = x_0 * w_8^j  (m = a + j n/8)
  constexpr bool PF = XFG_DEEP_PREFETCH && WC != 0;
  __shared__ u64 pf[PF ? WC + D : 1][PF ? DEEP_THREADS : 1];      // one buffer (static shared memory is capped at 48 KB): read into registers, then refilled
  auto fetch = [&](int j) {      // point j of this thread -> pf[.][tid]
    const size_t idx = (size_t)k * n + a + (size_t)j * n8;
#pragma unroll
    for (int c = 0; c < (PF ? WC : 0); c++) cp_async8(&pf[c][tid], lde + (size_t)c * N + idx);
#pragma unroll
    for (int l = 0; l < (PF ? D : 0); l++) cp_async8(&pf[(PF ? WC : 0) + l][tid], hlde + (size_t)l * N + idx);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (PF && active) fetch(0);
#pragma unroll DEEP_UNROLL
  for (int j = 0; j < (active ? 8 : 0); j++) {
    const size_t idx = (size_t)k * n + a + (size_t)j * n8;
    u64 pv[PF ? WC + D : 1];
    if (PF) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int c = 0; c < WC + D; c++) pv[c] = pf[c][tid];
      if (j + 1 < 8) fetch(j + 1);
    }
    DotAcc sa[D];     // S_T = sum_c gamma_c T_c(x): un-reduced dot product, one reduction per limb
#pragma unroll
    for (int c = 0; c < W; c++) { const u64 tv = PF ? pv[c] : lde[(size_t)c * N + idx]; for (int l = 0; l < D; l++) sa[l].fma(sc[2 * c + l], tv); }
    Ext<D> st; for (int l = 0; l < D; l++) st.set_limb(l, sa[l].result());
    Ext<D> h; for (int l = 0; l < D; l++) h.set_limb(l, PF ? pv[(PF ? WC : 0) + l] : hlde[(size_t)l * N + idx]);
    u64 uz, uzg;
    const u64 nz = deep_norm<D>(pz, x, uz), nzg = deep_norm<D>(pzg, x, uzg);
    const Ext<D> pc = deep_mul_conj<D>(pz, uz, st + delta * h - c1), qc = deep_mul_conj<D>(pzg, uzg, st - c2);
#pragma unroll
    for (int l = 0; l < D; l++) { DotAcc m; m.fma(pc.limb(l), nzg); m.fma(qc.limb(l), nz); sh[j][l][tid] = m.result(); }
    const u64 den = gl_mul(nz, nzg);
    sh[j][D][tid] = den; sh[j][D + 1][tid] = acc;
    acc = gl_mul(acc, den); x = gl_mul_pow2<24>(x);      // x_j = x_0 * w_8^j, w_8 = 2^24
  }
#if XFG_DEEP_CTA_INV
  __shared__ u64 invbuf[DEEP_THREADS];
  acc = cta_inv<DEEP_THREADS>(acc, invbuf);      // a thread without work hands in 1
  if (!active) return;
#else
  acc = w_inv(acc);
#endif
#pragma unroll 1
  for (int j = 7; j >= 0; j--) {
    const size_t idx = (size_t)k * n + a + (size_t)j * n8;
    const u64 dinv = gl_mul(sh[j][D + 1][tid], acc); acc = gl_mul(acc, sh[j][D][tid]);
#pragma unroll
    for (int l = 0; l < D; l++) { const u64 v = gl_mul(sh[j][l][tid], dinv); deep[(size_t)l * N + idx] = v; sh[j][l][tid] = v; }
  }
  if (fri_tree0) {
    u64 row[8 * D];
#pragma unroll
    for (int j = 0; j < 8; j++) for (int l = 0; l < D; l++) row[j * D + l] = sh[j][l][tid];
    store_digest(fri_tree0 + n + 8 * a + k, b3_hash_limbs<8 * D>(row));
  }
}
static size_t deep_smem(int D) { return (size_t)8 * (D + 2) * DEEP_THREADS * sizeof(u64); }
void launch_deep(cudaStream_t st, int D, const u64* lde, const u64* hlde, u32 ln, const ProofState* ps, const u64* dcoef, u32 width, PowTable wn, const u64* s_k,
                 u64* deep, Digest* fri_tree0) {
  const size_t n8 = (size_t(1) << ln) / 8; dim3 grid((unsigned)((n8 + DEEP_THREADS - 1) / DEEP_THREADS), 8);
  const u64 w8 = gl_root_of_unity(3);
  const size_t sm = deep_smem(D);
  if (width == XFG_TRACE_WIDTH) {
    if (D == 1) deep_kernel<1, XFG_TRACE_WIDTH><<<grid, DEEP_THREADS, sm, st>>>(lde, hlde, ln, ps, dcoef, width, wn, s_k, w8, deep, fri_tree0);
    else deep_kernel<2, XFG_TRACE_WIDTH><<<grid, DEEP_THREADS, sm, st>>>(lde, hlde, ln, ps, dcoef, width, wn, s_k, w8, deep, fri_tree0);
  } else if (D == 1) deep_kernel<1, 0><<<grid, DEEP_THREADS, sm, st>>>(lde, hlde, ln, ps, dcoef, width, wn, s_k, w8, deep, fri_tree0);
  else deep_kernel<2, 0><<<grid, DEEP_THREADS, sm, st>>>(lde, hlde, ln, ps, dcoef, width, wn, s_k, w8, deep, fri_tree0);
  XFG_LAUNCHED(1);
}
// Opt-in dynamic shared memory sizes are per device: called by every xfg_create for its own device (needed once XFG_DEEP_THREADS / XFG_CE_THREADS
// push the staging arrays past 48 KB; harmless below)
void stark_init() {
  cudaFuncSetAttribute(constraint_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ce_smem(1));
  cudaFuncSetAttribute(constraint_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ce_smem(2));
  cudaFuncSetAttribute(deep_kernel<1, XFG_TRACE_WIDTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)deep_smem(1));
  cudaFuncSetAttribute(deep_kernel<2, XFG_TRACE_WIDTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)deep_smem(2));
  cudaFuncSetAttribute(deep_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)deep_smem(1));
  cudaFuncSetAttribute(deep_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)deep_smem(2));
}

// ------------------------------------------------------------------------------------------------------------------
// FRI: fold layer l (Nl values) by 8 with alpha_l into layer l+1 and hash the rows of layer l+1 (A.10).
// Row r of layer l = values at r + j*Nl/8; x_r = 7 * w_Nl^r (constant offset 7 at every layer);
// next[r] = P_r(alpha), P_r interpolating the row over x_r * w_8^j:
//   P_r(alpha) = 1/8 sum_k V_k (alpha / x_r)^k,  V_k = sum_j v_j w_8^(-jk).
// Thread i' computes the 8 folds r = i' + q*Nl/64 = the 8 elements of row i' of layer l+1.
// src layout: coset-major when `src_coset` (layer 0 = DEEP evaluations), else natural [limb][i]; dst natural.
// ------------------------------------------------------------------------------------------------------------------

template <int D>
__global__ void __launch_bounds__(128) fri_fold_kernel(const u64* __restrict__ src, size_t src_limb_stride, int src_coset, u32 lNl, u32 layer,
                                                        const ProofState* __restrict__ ps, PowTable wN_inv, u32 lN, FriConsts fc,
                                                        u64* __restrict__ dst, size_t dst_limb_stride, Digest* __restrict__ next_tree) {
  const size_t Nl = size_t(1) << lNl, R = Nl / 8, Rn = R / 8;   // rows of this layer, rows of the next
  size_t tix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tix >= Rn) return;
  size_t ip = tix;
  if (src_coset && Rn >= 8) ip = ((tix % (Rn / 8)) << 3) | (tix / (Rn / 8));   // i' = 8a' + k with a' fastest: coalesced coset-major reads
  const Ext<D> alpha = ld_ext<D>(ps->alphas, layer);
  const size_t nrows0 = Nl / 8;   // coset stride n when src is the layer-0 coset-major array (Nl = 8n)
  __shared__ u64 sh[8 * D][128];  // the 8 folded values of this thread = one row of the next layer (rolled loop: the unrolled kernel was 300 KB of SASS)
#pragma unroll 1
  for (int q = 0; q < 8; q++) {
    const size_t r = ip + (size_t)q * Rn;
    Ext<D> v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const size_t p = r + (size_t)j * R;
      const size_t addr = src_coset ? (p & 7) * nrows0 + (p >> 3) : p;
      for (int l = 0; l < D; l++) v[j].set_limb(l, src[(size_t)l * src_limb_stride + addr]);
    }
    // 1 / x_r = 7^-1 * w_N^-(r * N/Nl)
    const u64 xinv = gl_mul(fc.inv7, pow_lookup(wN_inv, (u64)r << (lN - lNl)));
    Ext<D> w = fold8<D>(v, fc, mul_base(alpha, xinv));
#pragma unroll
    for (int l = 0; l < D; l++) { dst[(size_t)l * dst_limb_stride + r] = w.limb(l); sh[q * D + l][threadIdx.x] = w.limb(l); }
  }
  if (next_tree) {
    u64 row[8 * D];
#pragma unroll
    for (int i = 0; i < 8 * D; i++) row[i] = sh[i][threadIdx.x];
    store_digest(next_tree + Rn + ip, b3_hash_limbs<8 * D>(row));
  }
}
// Small layers (fewer next-layer rows than the GPU has thread slots): one fold per thread, the 8 threads of a next-layer row are
// adjacent lanes; lane 0 of the group hashes the row from shared memory.  Cuts the serial depth of a layer from 8 folds to 1.
template <int D>
__global__ void __launch_bounds__(128) fri_fold_small_kernel(const u64* __restrict__ src, size_t src_limb_stride, u32 lNl, u32 layer,
                                                              const ProofState* __restrict__ ps, PowTable wN_inv, u32 lN, FriConsts fc,
                                                              u64* __restrict__ dst, size_t dst_limb_stride, Digest* __restrict__ next_tree) {
  const size_t Nl = size_t(1) << lNl, R = Nl / 8, Rn = R / 8;
  const size_t tix = (size_t)blockIdx.x * blockDim.x + threadIdx.x, ip = tix >> 3; const u32 q = (u32)tix & 7, g = threadIdx.x >> 3;
  __shared__ u64 sh[16][8 * D];
  if (ip < Rn) {
    const size_t r = ip + (size_t)q * Rn;
    Ext<D> v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) for (int l = 0; l < D; l++) v[j].set_limb(l, src[(size_t)l * src_limb_stride + r + (size_t)j * R]);
    const u64 xinv = gl_mul(fc.inv7, pow_lookup(wN_inv, (u64)r << (lN - lNl)));
    const Ext<D> w = fold8<D>(v, fc, mul_base(ld_ext<D>(ps->alphas, layer), xinv));
#pragma unroll
    for (int l = 0; l < D; l++) { dst[(size_t)l * dst_limb_stride + r] = w.limb(l); sh[g][q * D + l] = w.limb(l); }
  }
  __syncthreads();
  if (next_tree && ip < Rn && q == 0) {
    u64 row[8 * D];
#pragma unroll
    for (int i = 0; i < 8 * D; i++) row[i] = sh[g][i];
    store_digest(next_tree + Rn + ip, b3_hash_limbs<8 * D>(row));
  }
}
void launch_fri_fold(cudaStream_t st, int D, const u64* src, size_t src_limb_stride, int src_coset, u32 lNl, u32 layer, const ProofState* ps,
                     PowTable wN_inv, u32 lN, const FriConsts& fc, u64* dst, size_t dst_limb_stride, Digest* next_tree) {
  const size_t Rn = (size_t(1) << lNl) / 64; const unsigned blocks = (unsigned)((Rn + 127) / 128);
  if (!src_coset && Rn <= 148 * 128 * 2) {
    const unsigned sb = (unsigned)((Rn * 8 + 127) / 128);
    if (D == 1) fri_fold_small_kernel<1><<<sb, 128, 0, st>>>(src, src_limb_stride, lNl, layer, ps, wN_inv, lN, fc, dst, dst_limb_stride, next_tree);
    else fri_fold_small_kernel<2><<<sb, 128, 0, st>>>(src, src_limb_stride, lNl, layer, ps, wN_inv, lN, fc, dst, dst_limb_stride, next_tree);
    XFG_LAUNCHED(1); return;
  }
  if (D == 1) fri_fold_kernel<1><<<blocks, 128, 0, st>>>(src, src_limb_stride, src_coset, lNl, layer, ps, wN_inv, lN, fc, dst, dst_limb_stride, next_tree);
  else fri_fold_kernel<2><<<blocks, 128, 0, st>>>(src, src_limb_stride, src_coset, lNl, layer, ps, wN_inv, lN, fc, dst, dst_limb_stride, next_tree);
  XFG_LAUNCHED(1);
}

// coset-major [k*n + m] -> natural [8m + k] (only used when a proof has no FRI layer and the remainder is taken from DEEP directly)
__global__ void coset_to_natural_kernel(const u64* __restrict__ src, u64* __restrict__ dst, u32 ln, int D, size_t src_limb_stride, size_t dst_limb_stride) {
  const size_t n = size_t(1) << ln, i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 8 * n) return;
  for (int l = 0; l < D; l++) dst[(size_t)l * dst_limb_stride + i] = src[(size_t)l * src_limb_stride + (i & 7) * n + (i >> 3)];
}
void launch_coset_to_natural(cudaStream_t st, const u64* src, u64* dst, u32 ln, int D, size_t src_limb_stride, size_t dst_limb_stride) {
  const size_t N = size_t(8) << ln;
  coset_to_natural_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(src, dst, ln, D, src_limb_stride, dst_limb_stride); XFG_LAUNCHED(1);
}

// build_trace on the device (src/burn_mint_air.rs:442-476 with the state column of SURVEY B.2): columns 0-3, 5, 6 are the constants the step-0
// assertions pin (AirParams::assert0), column 4 is floor(4 i / n).  The 8-argument entry point (xfg_prove_burn_mint_from_inputs) therefore
// uploads no trace at all.  Two u64 per thread, 16-byte stores.
__global__ void __launch_bounds__(256) trace_fill_kernel(u64* __restrict__ t, const AirParams* __restrict__ air, u32 ln) {
  const size_t half = size_t(1) << (ln - 1), i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const u32 c = blockIdx.y;
  ulonglong2 v;
  if (c == 4) { v.x = (8 * i) >> ln; v.y = (8 * i + 4) >> ln; } else { v.x = v.y = air->assert0[c]; }
  reinterpret_cast<ulonglong2*>(t + ((size_t)c << ln))[i] = v;
}
void launch_trace_fill(cudaStream_t st, u64* trace, const AirParams* d_air, u32 ln) {
  const size_t half = size_t(1) << (ln - 1);
  trace_fill_kernel<<<dim3((unsigned)((half + 255) / 256), XFG_TRACE_WIDTH), 256, 0, st>>>(trace, d_air, ln); XFG_LAUNCHED(1);
}

// ---- 32-bit integer-pipe peak (xfg_int_pipe_peak): 8 independent chains per thread of the BLAKE3 operation mix
// (LOP3 xor, IADD3, SHF rotate - all ALU-pipe instructions; 4 per group, checked in SASS), no memory traffic: the roofline denominator of the hashing kernels
__global__ void __launch_bounds__(256) int_peak_kernel(const u32* __restrict__ in, u32* __restrict__ out, u32 iters) {
  u32 x[8], y = in[threadIdx.x & 31], z = in[32 + (threadIdx.x & 31)];
#pragma unroll
  for (int c = 0; c < 8; c++) x[c] = in[(threadIdx.x + c) & 63];
  for (u32 i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int c = 0; c < 8; c++)   // xor, rotate, xor, rotate: 4 instructions that only the ALU pipe executes and ptxas cannot fuse or move (it turns
                                    // plain adds into IMAD on the FMA pipe and add-after-rotate into one LEA.HI, which would blur the count)
        asm volatile("xor.b32 %0, %0, %2;\n\t shf.r.wrap.b32 %0, %0, %0, 7;\n\t xor.b32 %0, %0, %1;\n\t shf.r.wrap.b32 %0, %0, %0, 11;" : "+r"(x[c]) : "r"(y), "r"(z));
  }
  u32 acc = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc ^= x[c];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// ---- pipe probe (xfg_pipe_probe): the same 8-chain loop with other instruction mixes, to see what the FMA pipe (IMAD) can take off the ALU pipe.
// MODE 1: 4 IMAD; 2: xor, IMAD, shf, IMAD; 3: xor, shf, xor, IMAD; 4: 4 IMAD.WIDE (64-bit accumulate); 5: xor, shf, IMAD.WIDE, xor; 6: 4 IADD3 (3-input adds)
template <int MODE> __global__ void __launch_bounds__(256) pipe_probe_kernel(const u32* __restrict__ in, u32* __restrict__ out, u32 iters) {
  u32 x[8], y = in[threadIdx.x & 31] | 1u, z = in[32 + (threadIdx.x & 31)];
  unsigned long long w[8];
#pragma unroll
  for (int c = 0; c < 8; c++) { x[c] = in[(threadIdx.x + c) & 63]; w[c] = x[c]; }
  for (u32 i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int c = 0; c < 8; c++) {
        if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;\n\t mad.lo.u32 %0, %0, %2, %1;\n\t mad.lo.u32 %0, %0, %1, %2;\n\t mad.lo.u32 %0, %0, %2, %1;" : "+r"(x[c]) : "r"(y), "r"(z));
        else if (MODE == 2) asm volatile("xor.b32 %0, %0, %2;\n\t mad.lo.u32 %0, %0, %1, %2;\n\t shf.r.wrap.b32 %0, %0, %0, 7;\n\t mad.lo.u32 %0, %0, %2, %1;" : "+r"(x[c]) : "r"(y), "r"(z));
        else if (MODE == 3) asm volatile("xor.b32 %0, %0, %2;\n\t shf.r.wrap.b32 %0, %0, %0, 7;\n\t xor.b32 %0, %0, %1;\n\t mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
        else if (MODE == 4) {       // w[c] += lo(w[c ^ 1]) * y : the multiplicand varies, nothing to hoist; 4 per group like the other modes
          asm volatile("{\n\t .reg .u32 lo, hi;\n\t mov.b64 {lo, hi}, %1;\n\t mad.wide.u32 %0, lo, %2, %0;\n\t mad.wide.u32 %0, hi, %3, %0;\n\t mad.wide.u32 %0, lo, %3, %0;\n\t mad.wide.u32 %0, hi, %2, %0;\n\t}"
                       : "+l"(w[c]) : "l"(w[c ^ 1]), "r"(y), "r"(z));
        }
        else if (MODE == 5) { asm volatile("xor.b32 %0, %0, %2;\n\t shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(x[c]) : "r"(y), "r"(z));
                              asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(x[c]), "r"(z));
                              asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[c]) : "r"(y)); }
        else if (MODE == 6) asm volatile("{\n\t .reg .u32 t;\n\t add.u32 t, %0, %1;\n\t add.u32 %0, t, %2;\n\t add.u32 t, %0, %2;\n\t add.u32 %0, t, %1;\n\t add.u32 t, %0, %1;\n\t add.u32 %0, t, %2;\n\t add.u32 t, %0, %2;\n\t add.u32 %0, t, %1;\n\t}" : "+r"(x[c]) : "r"(y), "r"(z));
      }
  }
  u32 acc = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc ^= x[c] ^ (u32)w[c] ^ (u32)(w[c] >> 32);
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
void launch_pipe_probe(cudaStream_t st, int mode, const u32* in, u32* out, u32 blocks, u32 iters) {
  switch (mode) {
    case 1: pipe_probe_kernel<1><<<blocks, 256, 0, st>>>(in, out, iters); break;
    case 2: pipe_probe_kernel<2><<<blocks, 256, 0, st>>>(in, out, iters); break;
    case 3: pipe_probe_kernel<3><<<blocks, 256, 0, st>>>(in, out, iters); break;
    case 4: pipe_probe_kernel<4><<<blocks, 256, 0, st>>>(in, out, iters); break;
    case 5: pipe_probe_kernel<5><<<blocks, 256, 0, st>>>(in, out, iters); break;
    case 6: pipe_probe_kernel<6><<<blocks, 256, 0, st>>>(in, out, iters); break;
    default: int_peak_kernel<<<blocks, 256, 0, st>>>(in, out, iters); break;
  }
  XFG_LAUNCHED(1);
}
void launch_int_peak(cudaStream_t st, const u32* in, u32* out, u32 blocks, u32 iters) { int_peak_kernel<<<blocks, 256, 0, st>>>(in, out, iters); XFG_LAUNCHED(1); }

// ---- field self-test (xfg_field_selftest): exercises the canonical and the weak arithmetic on caller-chosen operands ----
template <int S> __device__ __forceinline__ bool pow2_case(u32 op, u64 a, u64& r) { if (op == 100 + S) { r = w_canon(w_mul_pow2<S>(a)); return true; } return false; }
__global__ void field_selftest_kernel(u32 op, const u64* __restrict__ a, const u64* __restrict__ b, size_t n, u64* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 x = a[i], y = b[i]; u64 r = 0;
  switch (op) {
    case 0: r = gl_mul(x, y); break;                       // any u64 operands -> canonical
    case 1: r = w_canon(w_mul(x, y)); break;               // any u64 operands
    case 2: r = w_canon(w_add_c(x, y)); break;             // x weak, y canonical
    case 3: r = w_canon(w_sub_c(x, y)); break;             // x weak, y canonical
    case 4: r = gl_add(x, y); break;                       // canonical operands
    case 5: r = gl_sub(x, y); break;
    case 6: r = gl_inv(x); break;
    case 10: r = w_inv(x); break;                          // any u64 operand -> canonical
    case 7: r = w_canon(w_add_hi32(x, (u32)y)); break;
    case 8: r = w_canon(w_sub_hi32(x, (u32)y)); break;
    case 9: { DotAcc d; for (int i = 0; i < 37; i++) d.fma(x, y); d.fma(y, y); r = d.result(); } break;      // 37 x*y + y*y, any u64 operands
    default:
      pow2_case<0>(op, x, r) || pow2_case<1>(op, x, r) || pow2_case<12>(op, x, r) || pow2_case<24>(op, x, r) || pow2_case<31>(op, x, r) ||
      pow2_case<32>(op, x, r) || pow2_case<33>(op, x, r) || pow2_case<36>(op, x, r) || pow2_case<48>(op, x, r) || pow2_case<60>(op, x, r) ||
      pow2_case<63>(op, x, r) || pow2_case<64>(op, x, r) || pow2_case<65>(op, x, r) || pow2_case<72>(op, x, r) || pow2_case<84>(op, x, r) ||
      pow2_case<95>(op, x, r);
  }
  out[i] = r;
}
void launch_field_selftest(cudaStream_t st, u32 op, const u64* a, const u64* b, size_t n, u64* out) {
  field_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(op, a, b, n, out); XFG_LAUNCHED(1);
}

}  // namespace xfg
