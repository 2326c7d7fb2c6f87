// go_emul.cpp — TEST INFRASTRUCTURE.  Runs the general-options proof pipeline of the product (xfg-stark_b200/csrc/general_pipeline.cuh +
// general_bodies.cuh: the launch sequence and the per-thread kernel bodies) on the host: every body is called in a plain loop over its index
// space, the tuned NTT / Merkle kernels are replaced by textbook host loops.  This is how the pipeline is checked against the oracle and
// against the reference's own proofs on a CPU-only box, before (and independently of) the GPU parity tests.  Nothing in the product library
// links or loads this file; the product launches the same bodies as CUDA kernels only (general.cu) and has no CPU path.
#include <cstdio>
#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/asan_interface.h>
#endif
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "../../xfg-stark_b200/csrc/general_pipeline.cuh"
#include "../../xfg-stark_b200/csrc/air_compile.hpp"
#include "../../xfg-stark_b200/csrc/general_verify_host.hpp"

using namespace xfg;

namespace {

struct HostBK {
  template <class F> void run(size_t count, const F& f) { for (size_t t = 0; t < count; t++) f(t); }
  u64 grind_threads(u32) { return 1; }
  void merkle_upper(Digest* tree, size_t M) { for (size_t i = M - 1; i >= 1; i--) tree[i] = go_merge(tree[2 * i], tree[2 * i + 1]); }
  // NttJob semantics (ntt.cuh): transform t reads src + (t / src_div) * src_tstride, pre-scales element j by base_c^j (c = t % src_div),
  // transforms (natural order in and out), scales, post-scales output j by post_c^j (c = t % post_div), writes dst + t * dst_tstride
  void ntt(const NttJob& job) {
    const size_t n = size_t(1) << job.ln; const u32 sd = job.src_div ? job.src_div : 1, pd = job.post_div ? job.post_div : 1;
    u64 w = gl_root_of_unity(job.ln); if (job.inverse) w = gl_inv(w);
    std::vector<u64> tw(n / 2 ? n / 2 : 1); { u64 x = 1; for (size_t i = 0; i < n / 2; i++) { tw[i] = x; x = gl_mul(x, w); } }
    std::vector<u64> a(n);
    for (u32 t = 0; t < job.batch; t++) {
      const u64* src = job.src + (size_t)(t / sd) * job.src_tstride;
      for (size_t j = 0; j < n; j++) {
        u64 v = src[j];
        if (job.canon_flag && v >= GL_P) *job.canon_flag |= job.canon_bit;
        if (job.pre_lo) { const PowTable pt{job.pre_lo + (size_t)(t % sd) * POW_LO, job.pre_hi + (size_t)(t % sd) * job.pre_hi_stride}; v = gl_mul(v, pow_lookup(pt, j)); }
        a[j] = v;
      }
      // bit reversal + iterative radix-2 decimation in time
      for (size_t i = 1, j = 0; i < n; i++) { size_t bit = n >> 1; for (; j & bit; bit >>= 1) j ^= bit; j ^= bit; if (i < j) std::swap(a[i], a[j]); }
      for (size_t len = 2; len <= n; len <<= 1)
        for (size_t i = 0; i < n; i += len)
          for (size_t k = 0; k < len / 2; k++) { const u64 u = a[i + k], v = gl_mul(a[i + k + len / 2], tw[k * (n / len)]); a[i + k] = gl_add(u, v); a[i + k + len / 2] = gl_sub(u, v); }
      u64* dst = job.dst + (size_t)t * job.dst_tstride;
      for (size_t k = 0; k < n; k++) {
        u64 v = gl_mul(a[k], job.scale);
        if (job.post_lo) { const PowTable pt{job.post_lo + (size_t)(t % pd) * POW_LO, job.post_hi + (size_t)(t % pd) * job.post_hi_stride}; v = gl_mul(v, pow_lookup(pt, k)); }
        dst[k] = v;
      }
    }
  }
};

void set_err(char* err, size_t cap, const std::string& s) { if (err && cap) snprintf(err, cap, "%s", s.c_str()); }

int prove(const xfg_air_desc& air, const u64* trace, u32 n_log2, const uint32_t o6[6], u64 in_scale, u8* out, size_t cap, size_t* out_len, char* err, size_t errcap) {
  xfg_options o{}; o.num_queries = o6[0]; o.blowup_factor = o6[1]; o.grinding_factor = o6[2]; o.field_extension = o6[3]; o.fri_folding_factor = o6[4]; o.fri_remainder_max_degree = o6[5];
  const int D = (int)o.field_extension;
  std::unique_ptr<GenProgram> prog(new GenProgram); std::vector<u64> steps; std::string e;
  if (int rc = compile_air_impl(e, air, n_log2, *prog, steps)) { set_err(err, errcap, e); return rc; }
  const u32 K = go_comp_columns(prog->max_degree);
  GoPlan p;
  if (const char* why = go_plan_shape(p, n_log2, o.blowup_factor, o.fri_folding_factor, o.fri_remainder_max_degree, prog->max_degree)) { set_err(err, errcap, why); return XFG_ERR_BAD_OPTIONS; }
  std::vector<std::unique_ptr<std::vector<u64>>> keep;
  go_plan_tables(p, [&](const std::vector<u64>& v) { keep.emplace_back(new std::vector<u64>(v)); return (const u64*)keep.back()->data(); });
  for (size_t g = 0; g < steps.size(); g++) prog->group_point[g] = gl_pow(p.g_n, steps[g]);
  const u32 W = air.width;
  // under AddressSanitizer (tests/test_options_pins.py::test_emulated_pipeline_under_address_sanitizer) every region of the workspace is followed by a
  // poisoned guard zone, so an out-of-range index in any body - the same index arithmetic the CUDA kernels run - aborts the run
#if defined(__SANITIZE_ADDRESS__)
  const size_t gap = 64;
#else
  const size_t gap = 0;
#endif
  std::vector<std::pair<size_t, size_t>> gaps;
  GoCarve c; go_carve(nullptr, p, D, W, K, c, gap);
  std::vector<u64> slab(c.words, 0xA5A5A5A5A5A5A5A5ull); go_carve(slab.data(), p, D, W, K, c, gap, &gaps);
#if defined(__SANITIZE_ADDRESS__)
  for (auto& g : gaps) __asan_poison_memory_region(slab.data() + g.first, g.second * 8);
  struct Unpoison { std::vector<u64>& s; ~Unpoison() { __asan_unpoison_memory_region(s.data(), s.size() * 8); } } unpoison{slab};
  if (getenv("GO_EMUL_POKE")) c.trace_coef[((size_t)W * p.n + 7) & ~size_t(7)] = 1;      // self-test of the checker: one word past a region must abort
#endif
  std::memcpy(c.trace_in, trace, (size_t)W * p.n * 8);
  std::unique_ptr<GoState> st(new GoState); std::memset(st.get(), 0, sizeof(GoState));
  seed_elements(n_log2, o, W, air.pub_inputs, air.num_pub_inputs, st->seed_limbs);
  st->seed_count = 8 + air.num_pub_inputs; st->error_flags = 0; st->nonce = ~0ull;
  std::vector<GoGatherTask> tasks; const size_t mat_words = go_gather_tasks(p, D, W, K, o.num_queries, c, tasks);
  std::vector<u64> material(mat_words, 0);
  HostBK bk;
  go_enqueue(bk, D, p, c, st.get(), prog.get(), W, K, air.num_assertions, air.num_constraints + air.num_assertions, c.trace_in, in_scale, o.num_queries, o.grinding_factor, tasks, material.data());
  if (st->error_flags & ERR_FLAG_NONCANONICAL) { set_err(err, errcap, "non-canonical trace element"); return XFG_ERR_BAD_ARGS; }
  if (st->error_flags & ERR_FLAG_DEGREE) { set_err(err, errcap, "UnsatisfiedTransitionConstraintError"); return XFG_ERR_UNSATISFIED_CONSTRAINT; }
  if (st->error_flags & ERR_FLAG_COIN) { set_err(err, errcap, "FailedToDrawFieldElement"); return XFG_ERR_INTERNAL; }
  std::vector<u8> bytes; go_assemble(p, D, W, K, o, *st, material.data(), tasks, bytes);
  *out_len = bytes.size();
  if (bytes.size() > cap) { set_err(err, errcap, "output buffer too small"); return XFG_ERR_BUFFER_TOO_SMALL; }
  std::memcpy(out, bytes.data(), bytes.size());
  return XFG_OK;
}

}  // namespace

extern "C" {

// same flat AIR arrays as the oracle's orc_prove_air: desc = width, num_pub, num_const, num_instr, num_out, num_assert; code = 3 u32 per
// instruction (op, a, b); asr = 3 u64 per assertion (column, step, value)
int go_emul_prove_air(const uint32_t desc[6], const u64* pub, const u64* consts, const uint32_t* code, const uint32_t* outs, const u64* asr,
                      const u64* trace, uint32_t n_log2, const uint32_t o[6], u8* out, size_t cap, size_t* out_len, char* err, size_t errcap) {
  std::vector<xfg_air_instr> ins(desc[3]); for (uint32_t i = 0; i < desc[3]; i++) { ins[i].op = code[3 * i]; ins[i].a = code[3 * i + 1]; ins[i].b = code[3 * i + 2]; }
  std::vector<xfg_assertion> as(desc[5]); for (uint32_t i = 0; i < desc[5]; i++) { as[i].column = (uint32_t)asr[3 * i]; as[i].step = (uint32_t)asr[3 * i + 1]; as[i].value = asr[3 * i + 2]; }
  xfg_air_desc d{}; d.width = desc[0]; d.num_pub_inputs = desc[1]; d.num_constants = desc[2]; d.num_instr = desc[3]; d.num_constraints = desc[4]; d.num_assertions = desc[5];
  d.pub_inputs = pub; d.constants = consts; d.code = ins.data(); d.constraint_values = outs; d.assertions = as.data();
  return prove(d, trace, n_log2, o, 1, out, cap, out_len, err, errcap);
}
// verification of one proof by the product's general verifier body (general_verify.cuh) run on the host; returns the XFG_VERIFY_* code
int go_emul_verify_air(const uint32_t desc[6], const u64* pub, const u64* consts, const uint32_t* code, const uint32_t* outs, const u64* asr,
                       const u8* proof, size_t len, const uint32_t o6[6]) {
  std::vector<xfg_air_instr> ins(desc[3]); for (uint32_t i = 0; i < desc[3]; i++) { ins[i].op = code[3 * i]; ins[i].a = code[3 * i + 1]; ins[i].b = code[3 * i + 2]; }
  std::vector<xfg_assertion> as(desc[5]); for (uint32_t i = 0; i < desc[5]; i++) { as[i].column = (uint32_t)asr[3 * i]; as[i].step = (uint32_t)asr[3 * i + 1]; as[i].value = asr[3 * i + 2]; }
  xfg_air_desc d{}; d.width = desc[0]; d.num_pub_inputs = desc[1]; d.num_constants = desc[2]; d.num_instr = desc[3]; d.num_constraints = desc[4]; d.num_assertions = desc[5];
  d.pub_inputs = pub; d.constants = consts; d.code = ins.data(); d.constraint_values = outs; d.assertions = as.data();
  xfg_options o{}; o.num_queries = o6[0]; o.blowup_factor = o6[1]; o.grinding_factor = o6[2]; o.field_extension = o6[3]; o.fri_folding_factor = o6[4]; o.fri_remainder_max_degree = o6[5];
  std::vector<u8> bytes(((len + 7) & ~size_t(7)) + 8, 0); if (len) std::memcpy(bytes.data(), proof, len);
  GoVerifyRec rec; std::vector<u8> progs;
  go_verify_prepare(bytes.data(), len, d, o, 0, rec, progs);
  if (progs.empty()) progs.resize(8);
  std::unique_ptr<GoVerifyWork> work(new GoVerifyWork); int result = -1;
  HostBK bk;
  if (o.field_extension == 1) bk.run(1, GoVerify<1>{&rec, bytes.data(), progs.data(), work.get(), o, &result});
  else if (o.field_extension == 2) bk.run(1, GoVerify<2>{&rec, bytes.data(), progs.data(), work.get(), o, &result});
  else bk.run(1, GoVerify<3>{&rec, bytes.data(), progs.data(), work.get(), o, &result});
  return result;
}
// the burn-mint statement through the C++ AIR description the product uses for it (air_compile.hpp: BurnMintAirDesc); montgomery != 0: the trace
// is in Montgomery form (x * 2^64 mod p), undone by the interpolation's scale as in the product
int go_emul_prove_burn_mint(const u64 pi[12], const u64 consts[4], const u64* trace, uint32_t n_log2, const uint32_t o[6], int montgomery,
                            u8* out, size_t cap, size_t* out_len, char* err, size_t errcap) {
  xfg_air_consts air{}; for (int i = 0; i < XFG_NUM_PUB_INPUTS; i++) air.pub_inputs[i] = pi[i];
  air.txn_hash = consts[0]; air.recipient_hash = consts[1]; air.nullifier = consts[2]; air.commitment = consts[3];
  BurnMintAirDesc bm(air, n_log2);
  return prove(bm.d, trace, n_log2, o, montgomery ? gl_inv(0xFFFFFFFFull) : 1, out, cap, out_len, err, errcap);
}

}  // extern "C"
