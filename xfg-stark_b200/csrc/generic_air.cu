// generic_air.cu — the AIR-dependent kernels of the generic front-end for sm_100a.
//
// Replaces, for an AIR given as data (xfg_air_desc), what winter-prover 0.8.3 does around a user's `impl Air`:
// `Air::get_constraint_composition_coefficients`, `DefaultConstraintEvaluator::evaluate` calling `evaluate_transition`
// (e.g. src/winterfell_air.rs:87-113) over the constraint-evaluation domain, `BoundaryConstraints` built from
// `get_assertions` (src/winterfell_air.rs:117-124), and the channel steps that depend on the trace width
// (SURVEY.md A.8, A.9).  The program is interpreted: one thread per evaluation point runs the compiled register machine,
// control flow is uniform across the grid (every thread decodes the same instruction from a broadcast load), operands are
// coalesced loads of LDE columns, and the random linear combination is accumulated un-reduced (DotAcc).
#include "../../include/xfg_stark.h"
#include "generic_air.cuh"
#include "b3_rolled.cuh"
#include "field_weak.cuh"
#include "launch.cuh"

namespace xfg {

// (the trace-root step - coin seed, commit_trace, transition then boundary coefficients, A.8 - runs inside the tree kernel: merkle.cu RootStep)

// ------------------------------------------------------------------------------------------------------------------
// evaluate_constraints for a compiled program: thread t evaluates GCE_PTS points of constraint-evaluation coset k' (LDE coset 4k')
//   H(x) = T(x) (x - g^(n-1)) / (x^n - 1) + sum_groups B_g(x) / (x - g^step_g)                                 (A.8)
// The boundary sum is carried as one fraction num/den (num <- num (x - p_g) + B_g den, den <- den (x - p_g)), so a point costs
// one base-field inversion whatever the number of groups, batched over the thread's points.   out: [limb][k'][m], k' < 2
// ------------------------------------------------------------------------------------------------------------------
// GCE_PTS = points per thread: 4 for long traces (one inversion per 4 points), 1 when the domain is too small to fill the GPU otherwise.
static constexpr int GCE_THREADS = 128;
template <int D, int GCE_PTS>
__global__ void __launch_bounds__(GCE_THREADS) gen_constraint_kernel(const u64* __restrict__ lde, u32 ln, const GenProgram* __restrict__ prog, const GenState* __restrict__ gs,
                                                                     PowTable wn, u64 s_k0, u64 s_k1, u64 zinv0, u64 zinv1, u64 g_last, u64* __restrict__ out, size_t out_tstride) {
  __shared__ u64 sh[GCE_PTS][2 * D + 2][GCE_THREADS];      // per point: u (D), num (D), den, prefix
  __shared__ u64 sc[(GEN_MAX_CONSTRAINTS + GEN_MAX_ASSERTIONS) * 2];
  const size_t n = size_t(1) << ln, N = 8 * n;
  const u32 kp = blockIdx.y, k = kp * 4, tid = threadIdx.x;
  const size_t per = n / GCE_PTS, t = (size_t)blockIdx.x * blockDim.x + tid;
  const u32 T = prog->num_constraints, A = prog->num_assertions, G = prog->num_groups, NI = prog->num_instr;
  for (u32 i = tid; i < (T + A) * 2; i += blockDim.x) sc[i] = (&gs->coef[0][0])[i];
  __syncthreads();
  if (t >= per) return;
  const u64 sk = kp ? s_k1 : s_k0, zinv = kp ? zinv1 : zinv0;
  const u64* base = lde + (size_t)k * n;
  u64 acc = 1;
#pragma unroll 1
  for (int q = 0; q < GCE_PTS; q++) {
    const size_t m = t + q * per, mn = (m + 1) & (n - 1);
    u64 slot[GEN_MAX_SLOTS];
    auto fetch = [&](u32 kind, u32 idx) -> u64 {
      if (kind == GK_SLOT) return slot[idx];
      if (kind == GK_CONST) return prog->constants[idx];
      return base[(size_t)idx * N + (kind == GK_CUR ? m : mn)];      // frame = LDE rows 8m + k and 8(m+1) + k (A.8)
    };
    DotAcc ta[D];
#pragma unroll 1
    for (u32 i = 0; i < NI; i++) {
      const GenInstr in = prog->code[i];
      const u32 op = in.w0 & 15u, dst = in.w0 >> 8;
      const u64 va = fetch((in.w0 >> 4) & 3u, in.w1 & 0xFFFFu);
      if (op == GOP_OUT) {
#pragma unroll
        for (int l = 0; l < D; l++) ta[l].fma(sc[2 * dst + l], va);
        continue;
      }
      const u64 vb = fetch((in.w0 >> 6) & 3u, in.w1 >> 16);
      slot[dst] = op == GOP_MUL ? gl_mul(va, vb) : op == GOP_ADD ? gl_add(va, vb) : gl_sub(va, vb);
    }
    Ext<D> ts;
#pragma unroll
    for (int l = 0; l < D; l++) ts.set_limb(l, ta[l].result());
    const u64 x = gl_mul(sk, pow_lookup(wn, m));
    Ext<D> num; u64 den = 1; u32 ai = 0;
#pragma unroll 1
    for (u32 g = 0; g < G; g++) {
      DotAcc ba[D];
      for (; ai < A && prog->asr[ai].group == g; ai++) {
        const u64 dj = gl_sub(base[(size_t)prog->asr[ai].column * N + m], prog->asr[ai].value);
#pragma unroll
        for (int l = 0; l < D; l++) ba[l].fma(sc[2 * (T + ai) + l], dj);
      }
      Ext<D> bs;
#pragma unroll
      for (int l = 0; l < D; l++) bs.set_limb(l, ba[l].result());
      const u64 xg = gl_sub(x, prog->group_point[g]);
      num = mul_base(num, xg) + mul_base(bs, den);
      den = gl_mul(den, xg);
    }
    const Ext<D> u = mul_base(ts, gl_mul(gl_sub(x, g_last), zinv));
#pragma unroll
    for (int l = 0; l < D; l++) { sh[q][l][tid] = u.limb(l); sh[q][D + l][tid] = num.limb(l); }
    sh[q][2 * D][tid] = den; sh[q][2 * D + 1][tid] = acc;
    acc = gl_mul(acc, den);
  }
  acc = w_inv(acc);
#pragma unroll 1
  for (int q = GCE_PTS - 1; q >= 0; q--) {
    const size_t m = t + q * per;
    const u64 dinv = gl_mul(sh[q][2 * D + 1][tid], acc); acc = gl_mul(acc, sh[q][2 * D][tid]);
#pragma unroll
    for (int l = 0; l < D; l++) out[(size_t)(l * 2 + kp) * out_tstride + m] = gl_add(sh[q][l][tid], gl_mul(sh[q][D + l][tid], dinv));
  }
}
void launch_gen_constraints(cudaStream_t st, int D, const u64* lde, u32 ln, const GenProgram* prog, const GenState* gs, PowTable wn,
                            u64 s_k0, u64 s_k1, u64 zinv0, u64 zinv1, u64 g_last, u64* out, size_t out_tstride) {
  const int pts = ln >= 19 ? 4 : 1;
  const size_t per = (size_t(1) << ln) / pts; dim3 grid((unsigned)((per + GCE_THREADS - 1) / GCE_THREADS), 2);
  if (D == 1 && pts == 4) gen_constraint_kernel<1, 4><<<grid, GCE_THREADS, 0, st>>>(lde, ln, prog, gs, wn, s_k0, s_k1, zinv0, zinv1, g_last, out, out_tstride);
  else if (D == 1) gen_constraint_kernel<1, 1><<<grid, GCE_THREADS, 0, st>>>(lde, ln, prog, gs, wn, s_k0, s_k1, zinv0, zinv1, g_last, out, out_tstride);
  else if (pts == 4) gen_constraint_kernel<2, 4><<<grid, GCE_THREADS, 0, st>>>(lde, ln, prog, gs, wn, s_k0, s_k1, zinv0, zinv1, g_last, out, out_tstride);
  else gen_constraint_kernel<2, 1><<<grid, GCE_THREADS, 0, st>>>(lde, ln, prog, gs, wn, s_k0, s_k1, zinv0, zinv1, g_last, out, out_tstride);
  XFG_LAUNCHED(1);
}

// ------------------------------------------------------------------------------------------------------------------
// end of the out-of-domain step for a run-time width (A.9): sums the partials of the width + D polynomials, sends the interleaved
// frame and H(z) to the coin, draws width + 1 DEEP coefficients, and prepares C1 = sum gamma_j T_j(z) + delta H(z),
// C2 = sum gamma_j T_j(zg) for the DEEP kernel.  Warps share the partial sums; warp 0 then runs the transcript.
// ------------------------------------------------------------------------------------------------------------------
static constexpr int GOF_WARPS = 8;
template <int D> __global__ void __launch_bounds__(32 * GOF_WARPS) gen_ood_finish_kernel(ProofState* ps, GenState* gs, u32 W, const u64* __restrict__ partial, u32 nb) {
  __shared__ u64 sums[GEN_MAX_WIDTH + 2][2][2];
  __shared__ u64 limbs[2 * GEN_MAX_WIDTH * 2];
  const u32 lane = lane_id();
  for (u32 p = threadIdx.x >> 5; p < W + D; p += GOF_WARPS) {
    u64 s[4] = {0, 0, 0, 0};
    for (u32 b = lane; b < nb; b += 32)
#pragma unroll
      for (u32 q = 0; q < 4; q++) s[q] = gl_add(s[q], partial[((size_t)p * nb + b) * 4 + q]);
#pragma unroll
    for (u32 q = 0; q < 4; q++) {
      for (int o = 16; o > 0; o >>= 1) s[q] = gl_add(s[q], __shfl_xor_sync(0xFFFFFFFFu, s[q], o));
      if (lane == 0) sums[p][q >> 1][q & 1] = s[q];
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  Coin c = coin_load(ps);
  for (u32 t = lane; t < 2 * W * D; t += 32) { const u32 l = t % D, w = (t / D) & 1u, j = t / (2 * D); limbs[t] = sums[j][w][l]; }   // interleaved per column (A.9)
  __syncwarp();
  // hash_elements(frame): up to 4 BLAKE3 chunks of 128 limbs; lane c hashes chunk c, then every lane merges the chaining values (tree mode)
  Digest d;
  {
    const int nl = (int)(2 * W * D), chunks = nl <= 128 ? 1 : (nl + 127) / 128;
    Digest cv = iv_digest();
    { const int cc = (int)lane < chunks ? (int)lane : 0; int cl = nl - cc * 128; if (cl > 128) cl = 128; cv = r_chunk(limbs + cc * 128, cl, (u32)cc, chunks == 1); }
    const Digest c0 = bcast_digest(cv, 0), c1 = bcast_digest(cv, 1), c2 = bcast_digest(cv, 2), c3 = bcast_digest(cv, 3);
    if (chunks == 1) d = c0;
    else if (chunks == 2) d = r_parent(c0, c1, true);
    else { const Digest l = r_parent(c0, c1, false); d = chunks == 3 ? r_parent(l, c2, true) : r_parent(l, r_parent(c2, c3, false), true); }
  }
  r_reseed(c, d);
  // H(z) = P_limb0(z) + phi * P_limb1(z), phi = (0,1): (a0,a1) * phi = (-2 a1, a0 + a1)
  Ext<D> hz = ldx<D>(sums[W][0]);
  if (D == 2) { const u64 a0 = sums[W + 1][0][0], a1 = sums[W + 1][0][1]; hz = hz + Ext<D>(gl_neg(gl_dbl(a1)), gl_add(a0, a1)); }
  u64 hl[2] = {hz.limb(0), hz.limb(1)};
  r_reseed(c, r_hash_limbs(hl, D));
  const bool ok = r_draw_many<D>(c, W + 1, gs->dcoef);      // width trace coefficients, then 1 composition column
  __syncwarp();
  Ext<D> c1, c2;
  for (u32 j = lane; j < W; j += 32) { const Ext<D> g = ldx<D>(gs->dcoef[j]); c1 = c1 + g * ldx<D>(sums[j][0]); c2 = c2 + g * ldx<D>(sums[j][1]); }
  if (lane == 0) c1 = c1 + ldx<D>(gs->dcoef[W]) * hz;
  u64 r[4] = {c1.limb(0), D == 2 ? c1.limb(1) : 0, c2.limb(0), D == 2 ? c2.limb(1) : 0};
#pragma unroll
  for (int q = 0; q < 4; q++) for (int o = 16; o > 0; o >>= 1) r[q] = gl_add(r[q], __shfl_xor_sync(0xFFFFFFFFu, r[q], o));
  for (u32 t = lane; t < 2 * W; t += 32) { gs->ood_frame[t][0] = sums[t >> 1][t & 1][0]; gs->ood_frame[t][1] = D == 2 ? sums[t >> 1][t & 1][1] : 0; }
  if (lane == 0) {
    stx<D>(ps->hz, hz);
    if (!ok) ps->error_flags |= ERR_FLAG_COIN;
    ps->deep_c1[0] = r[0]; ps->deep_c1[1] = r[1]; ps->deep_c2[0] = r[2]; ps->deep_c2[1] = r[3];
  }
  coin_store(ps, c);
}
void launch_gen_ood_finish(cudaStream_t st, int D, ProofState* ps, GenState* gs, u32 width, const u64* partial, u32 nb) {
  if (D == 1) gen_ood_finish_kernel<1><<<1, 32 * GOF_WARPS, 0, st>>>(ps, gs, width, partial, nb); else gen_ood_finish_kernel<2><<<1, 32 * GOF_WARPS, 0, st>>>(ps, gs, width, partial, nb);
  XFG_LAUNCHED(1);
}

}  // namespace xfg
