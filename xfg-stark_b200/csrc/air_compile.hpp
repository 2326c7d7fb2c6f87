// air_compile.hpp — the host "compiler" of the generic AIR front-end: validates an xfg_air_desc the way winter-air 0.8.3 `AirContext::new` /
// `Assertion` / `BoundaryConstraints` would (A.3, A.8) and compiles its straight-line program to the register machine the constraint kernels
// interpret (generic_air.cu, general_bodies.cuh): dead-code elimination, liveness, slot allocation.  Plain C++ (no CUDA): shared by the
// product (air_api.inc, general_api.inc) and the CPU emulation harness of the general-options pipeline (tests/host_emul).
//
// Stands in for a user's `impl Air` (`evaluate_transition` + `get_assertions`, e.g. src/winterfell_air.rs:87-127, src/burn_mint_air.rs:335-395).
#pragma once
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/xfg_stark.h"
#include "generic_air.cuh"

namespace xfg {

// returns XFG_OK and fills `prog` (everything except group_point, which needs the trace length) or an error code
inline int compile_air_impl(std::string& err, const xfg_air_desc& d, u32 n_log2, GenProgram& prog, std::vector<u64>& group_steps) {
  auto fail = [&err](int code, const char* msg) { err = msg; return code; };
  const u32 w = d.width, C = d.num_constants, NI = d.num_instr, T = d.num_constraints, A = d.num_assertions;
  if (w < 1 || w > XFG_AIR_MAX_WIDTH) return fail(XFG_ERR_BAD_ARGS, "number of columns must be between 1 and 128 for this backend");
  if (d.num_pub_inputs > XFG_AIR_MAX_PUB_INPUTS || C > XFG_AIR_MAX_CONSTANTS || NI > XFG_AIR_MAX_INSTR) return fail(XFG_ERR_BAD_ARGS, "AIR description too large");
  if (T < 1) return fail(XFG_ERR_BAD_ARGS, "at least one transition constraint degree must be specified");
  if (A < 1) return fail(XFG_ERR_BAD_ARGS, "at least one assertion must be specified");
  if (T > XFG_AIR_MAX_CONSTRAINTS || A > XFG_AIR_MAX_ASSERTIONS) return fail(XFG_ERR_BAD_ARGS, "too many constraints or assertions");
  if ((d.num_pub_inputs && !d.pub_inputs) || (C && !d.constants) || (NI && !d.code) || !d.constraint_values || !d.assertions) return fail(XFG_ERR_BAD_ARGS, "null pointer in the AIR description");
  for (u32 i = 0; i < d.num_pub_inputs; i++) if (d.pub_inputs[i] >= XFG_P) return fail(XFG_ERR_BAD_ARGS, "non-canonical public input");
  for (u32 i = 0; i < C; i++) if (d.constants[i] >= XFG_P) return fail(XFG_ERR_BAD_ARGS, "non-canonical constant");
  const u32 first = 2 * w + C, total = first + NI;
  // degrees: frame values 1, constants 0, add/sub max, mul sum; winter-air derives ce_blowup and the number of composition columns
  // from the declared degrees (A.3) - degrees <= 2 keep both at the values this pipeline implements (2 and 1)
  std::vector<u32> deg(total, 0);
  for (u32 i = 0; i < 2 * w; i++) deg[i] = 1;
  for (u32 i = 0; i < NI; i++) {
    const xfg_air_instr& in = d.code[i];
    if (in.op > XFG_OP_MUL || in.a >= first + i || in.b >= first + i) return fail(XFG_ERR_BAD_ARGS, "invalid instruction");
    deg[first + i] = in.op == XFG_OP_MUL ? deg[in.a] + deg[in.b] : std::max(deg[in.a], deg[in.b]);
    if (deg[first + i] > XFG_AIR_MAX_DEGREE) return fail(XFG_ERR_UNSUPPORTED_OPTIONS, "transition constraint degree above 9 is not supported");
  }
  u32 max_degree = 1;
  for (u32 j = 0; j < T; j++) {
    if (d.constraint_values[j] >= total) return fail(XFG_ERR_BAD_ARGS, "invalid constraint output");
    if (deg[d.constraint_values[j]] == 0) return fail(XFG_ERR_BAD_ARGS, "transition constraint degree must be at least one");
    max_degree = std::max(max_degree, deg[d.constraint_values[j]]);
  }
  if ((u64)max_degree >= (u64(1) << n_log2)) return fail(XFG_ERR_BAD_ARGS, "transition constraint degree must be smaller than the trace length");
  prog.max_degree = max_degree; prog.pad_ = 0;
  // assertions in winter-air's order: (stride, first_step, column) = (step, column) for single assertions (A.8)
  std::vector<xfg_assertion> asr(d.assertions, d.assertions + A);
  std::sort(asr.begin(), asr.end(), [](const xfg_assertion& x, const xfg_assertion& y) { return x.step != y.step ? x.step < y.step : x.column < y.column; });
  group_steps.clear();
  for (u32 i = 0; i < A; i++) {
    if (asr[i].column >= w) return fail(XFG_ERR_BAD_ARGS, "assertion column out of range");
    if ((u64)asr[i].step >= (u64(1) << n_log2)) return fail(XFG_ERR_BAD_ARGS, "assertion step out of range");
    if (asr[i].value >= XFG_P) return fail(XFG_ERR_BAD_ARGS, "non-canonical assertion value");
    if (i && asr[i - 1].step == asr[i].step && asr[i - 1].column == asr[i].column) return fail(XFG_ERR_BAD_ARGS, "duplicate assertion");
    if (group_steps.empty() || group_steps.back() != asr[i].step) group_steps.push_back(asr[i].step);
    prog.asr[i] = GenAssertion{asr[i].column, (u32)group_steps.size() - 1, asr[i].value}; prog.asr_step[i] = asr[i].step;
  }
  if (group_steps.size() > XFG_AIR_MAX_GROUPS) return fail(XFG_ERR_UNSUPPORTED_OPTIONS, "too many distinct assertion steps for this backend");
  // liveness: instructions that (transitively) feed a constraint; last use of every instruction result
  std::vector<char> live(NI, 0); std::vector<int> last_use(NI, -1);
  for (u32 j = 0; j < T; j++) if (d.constraint_values[j] >= first) live[d.constraint_values[j] - first] = 1;
  for (int i = (int)NI - 1; i >= 0; i--) if (live[i]) { const xfg_air_instr& in = d.code[i]; if (in.a >= first) live[in.a - first] = 1; if (in.b >= first) live[in.b - first] = 1; }
  for (u32 i = 0; i < NI; i++) if (live[i]) { const xfg_air_instr& in = d.code[i]; if (in.a >= first) last_use[in.a - first] = (int)i; if (in.b >= first) last_use[in.b - first] = (int)i; }
  std::vector<std::vector<u32>> outs_of(NI);   // constraints whose value is instruction i
  std::vector<u32> direct;                     // constraints that are a frame value or a constant
  for (u32 j = 0; j < T; j++) { const u32 v = d.constraint_values[j]; if (v >= first) outs_of[v - first].push_back(j); else direct.push_back(j); }
  std::vector<int> slot_of(NI, -1); std::vector<u32> free_slots; u32 num_slots = 0, pc = 0;
  auto operand = [&](u32 v, u32& kind, u32& idx) {
    if (v < w) { kind = GK_CUR; idx = v; } else if (v < 2 * w) { kind = GK_NEXT; idx = v - w; } else if (v < first) { kind = GK_CONST; idx = v - 2 * w; }
    else { kind = GK_SLOT; idx = (u32)slot_of[v - first]; }
  };
  auto emit = [&](u32 op, u32 dst, u32 ak, u32 a, u32 bk, u32 b) { prog.code[pc++] = GenInstr{op | ak << 4 | bk << 6 | dst << 8, a | b << 16}; };
  for (u32 j : direct) { u32 k, x; operand(d.constraint_values[j], k, x); emit(GOP_OUT, j, k, x, 0, 0); }
  for (u32 i = 0; i < NI; i++) {
    if (!live[i]) continue;
    const xfg_air_instr& in = d.code[i];
    u32 ak, a, bk, b; operand(in.a, ak, a); operand(in.b, bk, b);
    // operands that die here give their slots back before the destination is chosen (the interpreter reads both operands first)
    for (u32 v : {in.a, in.b}) if (v >= first && last_use[v - first] == (int)i && slot_of[v - first] >= 0) { free_slots.push_back((u32)slot_of[v - first]); slot_of[v - first] = -1; }
    u32 dst;
    if (!free_slots.empty()) { dst = free_slots.back(); free_slots.pop_back(); } else dst = num_slots++;
    if (num_slots > XFG_AIR_MAX_LIVE) return fail(XFG_ERR_UNSUPPORTED_OPTIONS, "too many simultaneously live intermediate values");
    slot_of[i] = (int)dst;
    emit(in.op, dst, ak, a, bk, b);
    for (u32 j : outs_of[i]) emit(GOP_OUT, j, GK_SLOT, dst, 0, 0);
    if (last_use[i] < 0) { free_slots.push_back(dst); slot_of[i] = -1; }   // only constraints read it
  }
  prog.width = w; prog.num_constraints = T; prog.num_assertions = A; prog.num_groups = (u32)group_steps.size(); prog.num_instr = pc; prog.num_slots = num_slots;
  for (u32 i = 0; i < C; i++) prog.constants[i] = d.constants[i];
  return XFG_OK;
}

// highest transition-constraint degree of a description (0 if it is malformed - the compiler then reports why): decides the pipeline, since the tuned
// kernels assume one composition column (degree <= 2)
inline u32 air_max_degree(const xfg_air_desc& d) {
  const u32 w = d.width, first = 2 * w + d.num_constants;
  if (!d.code && d.num_instr) return 0;
  if (!d.constraint_values || w < 1 || w > XFG_AIR_MAX_WIDTH || d.num_instr > XFG_AIR_MAX_INSTR || d.num_constants > XFG_AIR_MAX_CONSTANTS) return 0;
  std::vector<u32> deg(first + d.num_instr, 0);
  for (u32 i = 0; i < 2 * w; i++) deg[i] = 1;
  for (u32 i = 0; i < d.num_instr; i++) {
    const xfg_air_instr& in = d.code[i];
    if (in.op > XFG_OP_MUL || in.a >= first + i || in.b >= first + i) return 0;
    deg[first + i] = std::min<u32>(64, in.op == XFG_OP_MUL ? deg[in.a] + deg[in.b] : std::max(deg[in.a], deg[in.b]));
  }
  u32 m = 1;
  for (u32 j = 0; j < d.num_constraints; j++) { if (d.constraint_values[j] >= deg.size()) return 0; m = std::max(m, deg[d.constraint_values[j]]); }
  return m;
}

// the normalised XfgBurnMintAir (src/burn_mint_air.rs:356-377 constraints, :383-394 assertions with the last step n - 1) as an AIR
// description: what the general-options pipeline proves when the burn-mint entry points are called with options outside the tuned 8/8 set.
// Same program as xfg-stark_b200/air.py: burn_mint_air (which the tuned kernels are checked against, tests/test_gpu_air.py).
struct BurnMintAirDesc {
  xfg_air_desc d{}; u64 consts[6]; xfg_air_instr code[11]; u32 outs[XFG_NUM_TRANSITION]; xfg_assertion asr[XFG_NUM_ASSERTIONS]; u64 pub[XFG_NUM_PUB_INPUTS];
  BurnMintAirDesc(const xfg_air_consts& air, u32 n_log2) {
    const u32 w = XFG_TRACE_WIDTH, C0 = 2 * w, I0 = C0 + 6;     // value ids: cur 0..6, next 7..13, constants 14..19, instructions 20..
    const u64 large = gl_mul(XFG_STD_BURN, 1000);
    consts[0] = XFG_STD_BURN; consts[1] = large; consts[2] = air.txn_hash; consts[3] = air.recipient_hash; consts[4] = air.nullifier; consts[5] = air.commitment;
    auto op = [](u32 o, u32 a, u32 b) { xfg_air_instr i; i.op = o; i.a = a; i.b = b; return i; };
    code[0] = op(XFG_OP_SUB, 0, C0 + 0); code[1] = op(XFG_OP_SUB, 0, C0 + 1); code[2] = op(XFG_OP_MUL, I0 + 0, I0 + 1);   // (c0 - std)(c0 - large)
    code[3] = op(XFG_OP_SUB, 1, 0);                                                                                    // c1 - c0
    code[4] = op(XFG_OP_SUB, 2, C0 + 2); code[5] = op(XFG_OP_SUB, 3, C0 + 3);                                          // c2 - txn, c3 - rcpt
    code[6] = op(XFG_OP_SUB, w + 4, 4);                                                                                // d = next4 - c4
    // d (d - 1) = d d - d: no constant 1 needed
    code[7] = op(XFG_OP_MUL, I0 + 6, I0 + 6); code[8] = op(XFG_OP_SUB, I0 + 7, I0 + 6);
    code[9] = op(XFG_OP_SUB, 5, C0 + 4); code[10] = op(XFG_OP_SUB, 6, C0 + 5);                                         // c5 - nullifier, c6 - commitment
    outs[0] = I0 + 2; outs[1] = I0 + 3; outs[2] = I0 + 4; outs[3] = I0 + 5; outs[4] = I0 + 8; outs[5] = I0 + 9; outs[6] = I0 + 10;
    for (int i = 0; i < XFG_NUM_PUB_INPUTS; i++) pub[i] = air.pub_inputs[i];
    const u64 a0[XFG_TRACE_WIDTH] = {air.pub_inputs[XFG_PI_BURN], air.pub_inputs[XFG_PI_MINT], air.pub_inputs[XFG_PI_TXN_HASH], air.pub_inputs[XFG_PI_RECIPIENT_HASH], 0, air.nullifier, air.commitment};
    for (u32 j = 0; j < w; j++) { asr[j].column = j; asr[j].step = 0; asr[j].value = a0[j]; }
    asr[w].column = 4; asr[w].step = (u64(1) << n_log2) - 1; asr[w].value = XFG_FINAL_STATE;
    d.width = w; d.num_pub_inputs = XFG_NUM_PUB_INPUTS; d.num_constants = 6; d.num_instr = 11; d.num_constraints = XFG_NUM_TRANSITION; d.num_assertions = XFG_NUM_ASSERTIONS;
    d.pub_inputs = pub; d.constants = consts; d.code = code; d.constraint_values = outs; d.assertions = asr;
  }
};

}  // namespace xfg
