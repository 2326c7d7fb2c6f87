// general.cuh — launcher of the general-options proof pipeline on a CUDA stream (see general_pipeline.cuh / general_bodies.cuh).
#pragma once
#include <cuda_runtime.h>
#include "general_pipeline.cuh"
#include "general_verify.cuh"

namespace xfg {

// enqueues the whole proof (extend_execution_trace .. build_proof_object) on `st`; no host synchronisation
void go_launch(cudaStream_t st, int D, const GoPlan& p, const GoCarve& c, GoState* s, const GenProgram* prog, u32 W, u32 K, u32 num_assertions, u32 ncoef, const u64* trace_src, u64 in_scale,
               u32 num_queries, u32 grinding, const std::vector<GoGatherTask>& tasks, u64* material);

// one thread per proof (general_verify.cuh); results[i] = XFG_VERIFY_*
void go_verify_launch(cudaStream_t st, int D, u32 count, const GoVerifyRec* recs, const u8* bytes, const u8* progs, GoVerifyWork* work, const xfg_options& opt, int* results);

}  // namespace xfg
