// generic_air.cuh — device-side representation of a generic AIR (the kernels launched from here serve degree <= 2, general_bodies.cuh the rest) and the launchers of the
// kernels that depend on the AIR: coefficient draws, constraint evaluation (a register-machine interpreter), the end of the
// out-of-domain step and the DEEP composition for a run-time trace width.
//
// Plays the role of a user's `impl Air` (`evaluate_transition` + `get_assertions`, e.g. src/winterfell_air.rs:87-127) inside
// winter-prover 0.8.3 `DefaultConstraintEvaluator` / `BoundaryConstraints` (A.8).  Everything that does not depend on the AIR
// (NTTs, row hashing, Merkle trees, FRI, grinding, queries) is the burn-mint pipeline's code, unchanged.
#pragma once
#include <cuda_runtime.h>
#include "state.cuh"
#include "field.cuh"

namespace xfg {

static constexpr int GEN_MAX_WIDTH = XFG_AIR_MAX_WIDTH, GEN_MAX_CONSTS = XFG_AIR_MAX_CONSTANTS, GEN_MAX_INSTR = XFG_AIR_MAX_INSTR + XFG_AIR_MAX_CONSTRAINTS,
                     GEN_MAX_CONSTRAINTS = XFG_AIR_MAX_CONSTRAINTS, GEN_MAX_ASSERTIONS = XFG_AIR_MAX_ASSERTIONS, GEN_MAX_GROUPS = XFG_AIR_MAX_GROUPS,
                     GEN_MAX_SLOTS = XFG_AIR_MAX_LIVE;

// operand kinds of a compiled instruction
enum : u32 { GK_CUR = 0, GK_NEXT = 1, GK_CONST = 2, GK_SLOT = 3 };
// compiled opcodes: the three field operations write slot `dst`; OUT adds coef[dst] * operand a to the transition combination
enum : u32 { GOP_ADD = 0, GOP_SUB = 1, GOP_MUL = 2, GOP_OUT = 3 };
// word0 = op | akind << 4 | bkind << 6 | dst << 8 ; word1 = a | b << 16
struct GenInstr { u32 w0, w1; };

struct GenAssertion { u32 column, group; u64 value; };

// the compiled AIR, one copy per proof slot in device memory (written by the host through a pinned mirror)
struct GenProgram {
  u32 width, num_constraints, num_assertions, num_groups, num_instr, num_slots;
  u32 max_degree, pad_;                     // highest transition-constraint degree: > 2 means several composition columns (general-options pipeline only)
  u64 group_point[GEN_MAX_GROUPS];          // g^step of each boundary-constraint group: the root of its divisor (x - g^step)
  u64 constants[GEN_MAX_CONSTS];
  GenAssertion asr[GEN_MAX_ASSERTIONS + 1]; // sorted (step, column): groups are contiguous runs
  u32 asr_step[GEN_MAX_ASSERTIONS + 1];     // step of each assertion (trace validation of the general-options pipeline)
  GenInstr code[GEN_MAX_INSTR];
};

// AIR-sized parts of the proof state (the fixed-size ProofState keeps the coin, roots, z, H(z), alphas, remainder, queries)
struct GenState {
  u64 ood_frame[2 * GEN_MAX_WIDTH][2];                          // T_0(z), T_0(zg), T_1(z), ... (A.9 interleaving); copied back to the host
  u64 dcoef[GEN_MAX_WIDTH + 1][2];                              // DEEP coefficients: width trace columns, then the composition column
  u64 coef[GEN_MAX_CONSTRAINTS + GEN_MAX_ASSERTIONS][2];        // transition coefficients, then boundary coefficients (A.8 draw order)
};

void launch_gen_constraints(cudaStream_t st, int D, const u64* lde, u32 ln, const GenProgram* prog, const GenState* gs, PowTable wn,
                            u64 s_k0, u64 s_k1, u64 zinv0, u64 zinv1, u64 g_last, u64* out, size_t out_tstride);
void launch_gen_ood_finish(cudaStream_t st, int D, ProofState* ps, GenState* gs, u32 width, const u64* partial, u32 nb);

}  // namespace xfg
