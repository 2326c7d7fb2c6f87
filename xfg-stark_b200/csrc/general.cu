// general.cu — the general-options proof pipeline as CUDA kernels for sm_100a: every body of general_bodies.cuh is launched through one
// generic kernel (one thread per output value); interpolation / LDE and the Merkle levels above the leaves are the tuned kernels of
// ntt.cu / merkle.cu.
//
// Replaces winter-prover 0.8.3 `Prover::prove` (src/burn_mint_prover.rs:124) for the `ProofOptions` that `XfgBurnMintProver::with_options`
// (src/burn_mint_prover.rs:44-49) accepts and the tuned pipeline of prover.cu does not serve: blowup factor != 8, FRI folding factor != 8,
// remainder degree < 7, FieldExtension::Cubic.
#include "general.cuh"
#include "merkle.cuh"
#include "launch.cuh"

namespace xfg {

static constexpr int GO_THREADS = 128;
template <class F> __global__ void __launch_bounds__(GO_THREADS) go_kernel(size_t count, const F f) {
  const size_t t = (size_t)blockIdx.x * GO_THREADS + threadIdx.x;
  if (t < count) f(t);
}

namespace {
struct GpuBK {
  cudaStream_t st; const GoPlan* plan;
  template <class F> void run(size_t count, const F& f) {
    if (!count) return;
    go_kernel<F><<<(unsigned)((count + GO_THREADS - 1) / GO_THREADS), GO_THREADS, 0, st>>>(count, f); XFG_LAUNCHED(1);
  }
  void ntt(const NttJob& j) { ntt_batch(st, plan->ntt, j); }
  void merkle_upper(Digest* tree, size_t M) { merkle_build_upper(st, tree, M, nullptr); }
  // small grinding factors need a handful of candidates; large ones use the whole chip (as launch_grind, transcript.cu)
  u64 grind_threads(u32 grinding) { return grinding <= 8 ? 1024 : grinding <= 16 ? 148 * 256 : 148 * 8 * 256; }
};
}  // namespace

void go_launch(cudaStream_t st, int D, const GoPlan& p, const GoCarve& c, GoState* s, const GenProgram* prog, u32 W, u32 K, u32 num_assertions, u32 ncoef, const u64* trace_src, u64 in_scale,
               u32 num_queries, u32 grinding, const std::vector<GoGatherTask>& tasks, u64* material) {
  GpuBK bk{st, &p};
  go_enqueue(bk, D, p, c, s, prog, W, K, num_assertions, ncoef, trace_src, in_scale, num_queries, grinding, tasks, material);
}

void go_verify_launch(cudaStream_t st, int D, u32 count, const GoVerifyRec* recs, const u8* bytes, const u8* progs, GoVerifyWork* work, const xfg_options& opt, int* results) {
  GpuBK bk{st, nullptr};
  if (D == 1) bk.run(count, GoVerify<1>{recs, bytes, progs, work, opt, results});
  else if (D == 2) bk.run(count, GoVerify<2>{recs, bytes, progs, work, opt, results});
  else bk.run(count, GoVerify<3>{recs, bytes, progs, work, opt, results});
}

}  // namespace xfg
