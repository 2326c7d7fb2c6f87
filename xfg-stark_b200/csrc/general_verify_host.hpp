// general_verify_host.hpp — host preparation of a general-options verification batch (general_verify.cuh): per proof the trace length is read from
// the proof's context bytes, the AIR description is validated and compiled for that length (air_compile.hpp), and the coin seed elements are laid out
// (Context::to_elements || public inputs, A.4).  Plain C++: shared by the product (general_api.inc) and tests/host_emul.
#pragma once
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "air_compile.hpp"
#include "general_verify.cuh"
#include "proof_bytes.hpp"

namespace xfg {

// appends the compiled program of `air` to `progs` (8-byte aligned) and fills `rec`; a proof that cannot even name its trace length, or an AIR that
// does not validate for it, is rejected here (rec.host_status) and costs no device work
inline void go_verify_prepare(const u8* proof, size_t len, const xfg_air_desc& air, const xfg_options& o, u64 proof_off, GoVerifyRec& rec, std::vector<u8>& progs) {
  std::memset(&rec, 0, sizeof rec);
  rec.proof_off = proof_off; rec.proof_len = (u32)len;
  if (!proof || len < 4 || len >= (size_t(1) << 32) || proof[3] < 3 || proof[3] > 27) { rec.host_status = XFG_VERIFY_MALFORMED; return; }
  const u32 n_log2 = proof[3];
  std::unique_ptr<GenProgram> prog(new GenProgram); std::vector<u64> steps; std::string err;
  if (compile_air_impl(err, air, n_log2, *prog, steps)) { rec.host_status = XFG_VERIFY_MALFORMED; return; }
  if (air.num_pub_inputs + 8 > (u32)MAX_SEED_LIMBS) { rec.host_status = XFG_VERIFY_MALFORMED; return; }
  const u64 g_n = gl_root_of_unity(n_log2);
  for (size_t g = 0; g < steps.size(); g++) prog->group_point[g] = gl_pow(g_n, steps[g]);
  seed_elements(n_log2, o, air.width, air.pub_inputs, air.num_pub_inputs, rec.seed_limbs);
  rec.seed_count = 8 + air.num_pub_inputs;
  const size_t used = offsetof(GenProgram, code) + (size_t)prog->num_instr * sizeof(GenInstr);
  rec.prog_off = (u32)progs.size();
  progs.resize(progs.size() + ((used + 7) & ~size_t(7)));
  std::memcpy(progs.data() + rec.prog_off, prog.get(), used);
}

}  // namespace xfg
