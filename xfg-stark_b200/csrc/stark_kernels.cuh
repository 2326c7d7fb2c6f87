// stark_kernels.cuh — launchers of the elementwise proof stages (see stark_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "state.cuh"
#if defined(__CUDACC__)
#include "field_weak.cuh"
#endif

namespace xfg {

struct FriConsts { u64 w8i[4]; u64 inv8; u64 inv7; };   // w_8^-1 powers 0..3, 8^-1, 7^-1
#if defined(__CUDACC__)
// x * 2^S (S < 96) and x * (-2^S) for a canonical x, canonical result: shifts and carry fix-ups instead of a 64x64-bit multiplication
template <int S> __device__ __forceinline__ u64 gl_mul_pow2(u64 x) { return w_canon(w_mul_pow2<S>(x)); }
template <int S, int D> __device__ __forceinline__ Ext<D> ext_mul_neg_pow2(Ext<D> a) {
  Ext<D> r;
#pragma unroll
  for (int l = 0; l < D; l++) r.set_limb(l, gl_neg(gl_mul_pow2<S>(a.limb(l))));
  return r;
}
// P(beta * x_r) for the polynomial P of degree < 8 interpolating v[j] at x_r * w_8^j, i.e. apply_drp's fold with beta = alpha / x_r
// (shared by the prover's FRI layers and the verifier's per-query fold check)
template <int D>
__device__ __forceinline__ Ext<D> fold8(const Ext<D> (&v)[8], const FriConsts& fc, Ext<D> beta) {
  // radix-2 DIT inverse DFT of size 8 (input bit-reversed), twiddles w_8^-j.  w_8 = 2^24 and 2^96 = -1, so w_8^-j = -2^(96 - 24 j): powers of two
  Ext<D> a[8] = {v[0], v[4], v[2], v[6], v[1], v[5], v[3], v[7]};
#pragma unroll
  for (int i = 0; i < 8; i += 2) { Ext<D> u = a[i], w = a[i + 1]; a[i] = u + w; a[i + 1] = u - w; }
#pragma unroll
  for (int i = 0; i < 8; i += 4) {
    Ext<D> u = a[i], w = a[i + 2]; a[i] = u + w; a[i + 2] = u - w;
    u = a[i + 1]; w = ext_mul_neg_pow2<48, D>(a[i + 3]); a[i + 1] = u + w; a[i + 3] = u - w;      // w_8^-2 = -2^48
  }
  { Ext<D> u = a[0], w = a[4]; a[0] = u + w; a[4] = u - w; }
  { Ext<D> u = a[1], w = ext_mul_neg_pow2<72, D>(a[5]); a[1] = u + w; a[5] = u - w; }               // w_8^-1 = -2^72
  { Ext<D> u = a[2], w = ext_mul_neg_pow2<48, D>(a[6]); a[2] = u + w; a[6] = u - w; }               // w_8^-2
  { Ext<D> u = a[3], w = ext_mul_neg_pow2<24, D>(a[7]); a[3] = u + w; a[7] = u - w; }               // w_8^-3 = -2^24
  Ext<D> r = a[7];
#pragma unroll
  for (int kk = 6; kk >= 0; kk--) r = r * beta + a[kk];
  return mul_base(r, fc.inv8);
}
#endif
void launch_constraints(cudaStream_t st, int D, const u64* lde, u32 ln, const AirParams* d_air, const ProofState* ps, PowTable wn,
                        u64 s_k0, u64 s_k1, u64 zinv0, u64 zinv1, u64* out, size_t out_tstride);   // evaluation (limb l, coset k') at out + (2l + k') * out_tstride
void launch_combine(cudaStream_t st, const u64* a, u32 ln, int D, u64 inv2, u64* h, ProofState* ps);
u32 ood_num_blocks(u32 ln);
void launch_ood(cudaStream_t st, int D, const u64* trace_coef, const u64* h_coef, u32 ln, u32 width, const ProofState* ps, u64* partial);
void stark_init();   // per-device kernel attributes (dynamic shared memory opt-in); xfg_create calls it
void launch_deep(cudaStream_t st, int D, const u64* lde, const u64* hlde, u32 ln, const ProofState* ps, const u64* dcoef, u32 width, PowTable wn, const u64* s_k,
                 u64* deep, Digest* fri_tree0);
void launch_fri_fold(cudaStream_t st, int D, const u64* src, size_t src_limb_stride, int src_coset, u32 lNl, u32 layer, const ProofState* ps,
                     PowTable wN_inv, u32 lN, const FriConsts& fc, u64* dst, size_t dst_limb_stride, Digest* next_tree);
void launch_int_peak(cudaStream_t st, const u32* in, u32* out, u32 blocks, u32 iters);
void launch_pipe_probe(cudaStream_t st, int mode, const u32* in, u32* out, u32 blocks, u32 iters);
void launch_field_selftest(cudaStream_t st, u32 op, const u64* a, const u64* b, size_t n, u64* out);
// fri_tail.cu: the FRI layers of at most 2^FRI_TAIL_MAX_LOG evaluations, the remainder, the grinding nonce and the query positions in one launch
// Measured (B200, round 2; ms per proof at 2^20 rows / quadratic, 2^16 rows / no extension, 1024 x 2^16 proofs/s): 14 -> 3.880, 0.4855, 4699;
// 13 -> 3.855, 0.4852, 4766; 12 -> 3.856, 0.482, 4725; 11 -> 3.854, 0.483, 4775.  At 14 the 2^14-evaluation layer of a 2^20-row proof (tree 20 us + fold
// 28 us on ONE SM) sits in the tail; as two multi-CTA launches it costs half.  Below 13 the gain is within the noise and every layer adds two launches.
#ifndef XFG_FRI_TAIL_MAX_LOG
#define XFG_FRI_TAIL_MAX_LOG 13
#endif
static constexpr u32 FRI_TAIL_MAX_LOG = XFG_FRI_TAIL_MAX_LOG, FRI_TAIL_MAX_GRIND = 12, NTT_TW_LOG_TAIL = 12;
struct FriTailArgs {
  u32 first_layer, num_layers, layer_log[MAX_LAYERS + 1];      // layers [first_layer, num_layers) are folded here
  u64* evals[MAX_LAYERS + 1]; u64 limb_stride[MAX_LAYERS + 1]; // evaluations of layer l: natural order [limb][i] (layer 0: coset-major DEEP evaluations)
  Digest* tree[MAX_LAYERS];                                    // heap-ordered trees; the leaves of tree[first_layer] are already hashed
  const u64* rem_in; u64 rem_stride; u32 rem_log, rem_len; u64 rem_ninv;   // remainder: evaluations of the last layer, natural order
  const u64* tw_inv; const u64* un_lo;                         // w_4096^-i (i < 2048); 7^-j (j < 4096)
  PowTable wN_inv; u32 lN; FriConsts fc;
  u32 grinding, num_queries, do_grind;                         // do_grind = 0: stop after the remainder (grind / positions kernels follow)
  u32 limbs_off;                                               // set by the launcher: shared-memory offset (words) of the remainder limbs
};
void launch_fri_tail(cudaStream_t st, int D, FriTailArgs a, ProofState* ps, u32 threads);
void launch_trace_fill(cudaStream_t st, u64* trace, const AirParams* d_air, u32 ln);   // burn-mint build_trace on the device
void launch_coset_to_natural(cudaStream_t st, const u64* src, u64* dst, u32 ln, int D, size_t src_limb_stride, size_t dst_limb_stride);

}  // namespace xfg
