// oracle/capi.cpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// C entry points over the oracle headers so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
// call it through ctypes.  Nothing in xfg-stark_b200/ includes, links or loads this library.
#include <cstdio>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "prover.hpp"
#include "verifier.hpp"

namespace orc { int g_threads = 1; }
using namespace orc;

namespace {
ProverDebug<F1> g_dbg1; ProverDebug<F2> g_dbg2; ProverDebug<F3> g_dbg3; int g_dbg_ext = 0;
ProofOptions opts_from(const uint32_t o[6]) { ProofOptions p; p.num_queries = o[0]; p.blowup = o[1]; p.grinding = o[2]; p.ext = o[3]; p.folding = o[4]; p.rem_max_deg = o[5]; return p; }
PublicInputs pi_from(const u64* v) { PublicInputs p; for (int i = 0; i < XFG_NUM_PUB_INPUTS; i++) p.v[i] = v[i]; return p; }
AirConsts ac_from(const u64* c) { return {c[0], c[1], c[2], c[3]}; }
template <class E> void to_limbs(const std::vector<E>& v, std::vector<u64>& o) { o.clear(); for (auto& e : v) for (int l = 0; l < E::DEG; l++) o.push_back(e.limb(l)); }
void set_err(char* err, size_t cap, const std::string& s) { if (err && cap) { snprintf(err, cap, "%s", s.c_str()); } }
}  // namespace

extern "C" {

int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_threads(int t) { g_threads = t < 1 ? 1 : t; }

u64 orc_fadd(u64 a, u64 b) { return fadd(a, b); }
u64 orc_fsub(u64 a, u64 b) { return fsub(a, b); }
u64 orc_fmul(u64 a, u64 b) { return fmul(a, b); }
u64 orc_fmul_slow(u64 a, u64 b) { return fmul_slow(a, b); }
u64 orc_finv(u64 a) { return finv(a); }
u64 orc_fpow(u64 a, u64 e) { return fpow(a, e); }
u64 orc_root_of_unity(unsigned k) { return root_of_unity(k); }
void orc_f2_mul(const u64 a[2], const u64 b[2], u64 o[2]) { F2 r = F2(a[0], a[1]) * F2(b[0], b[1]); o[0] = r.a0; o[1] = r.a1; }
void orc_f2_inv(const u64 a[2], u64 o[2]) { F2 r = F2(a[0], a[1]).inv(); o[0] = r.a0; o[1] = r.a1; }
void orc_f3_mul(const u64 a[3], const u64 b[3], u64 o[3]) { F3 r = F3(a[0], a[1], a[2]) * F3(b[0], b[1], b[2]); for (int i = 0; i < 3; i++) o[i] = r.a[i]; }
void orc_f3_inv(const u64 a[3], u64 o[3]) { F3 r = F3(a[0], a[1], a[2]).inv(); for (int i = 0; i < 3; i++) o[i] = r.a[i]; }

// data: n elements of `deg` limbs each (interleaved limbs). mode 0: forward NTT; 1: interpolate_poly (scaled inverse);
// 2: naive O(n^2) forward DFT
void orc_ntt(u64* data, size_t n, int deg, int mode) {
  if (deg == 1) {
    std::vector<F1> v(n); for (size_t i = 0; i < n; i++) v[i] = F1(data[i]);
    if (mode == 0) ntt(v); else if (mode == 1) interpolate_poly(v); else v = naive_dft(v, root_of_unity(ilog2(n)));
    for (size_t i = 0; i < n; i++) data[i] = v[i].v;
  } else {
    std::vector<F2> v(n); for (size_t i = 0; i < n; i++) v[i] = F2(data[2 * i], data[2 * i + 1]);
    if (mode == 0) ntt(v); else if (mode == 1) interpolate_poly(v); else v = naive_dft(v, root_of_unity(ilog2(n)));
    for (size_t i = 0; i < n; i++) { data[2 * i] = v[i].a0; data[2 * i + 1] = v[i].a1; }
  }
}
// evaluate_poly_with_offset over the base field: out[i] = p(offset * w_N^i), N = n * blowup
void orc_lde(const u64* coeffs, size_t n, size_t blowup, u64 offset, u64* out) {
  std::vector<F1> p(n); for (size_t i = 0; i < n; i++) p[i] = F1(coeffs[i]);
  std::vector<F1> r = evaluate_poly_with_offset(p, offset, blowup);
  for (size_t i = 0; i < r.size(); i++) out[i] = r[i].v;
}
void orc_interpolate_offset(u64* data, size_t n, u64 offset) {
  std::vector<F1> v(n); for (size_t i = 0; i < n; i++) v[i] = F1(data[i]);
  interpolate_poly_with_offset(v, offset);
  for (size_t i = 0; i < n; i++) data[i] = v[i].v;
}
void orc_blake3(const u8* p, size_t len, u8 out[32]) { Digest d = blake3(p, len); std::memcpy(out, d.data(), 32); }
void orc_keccak256(const u8* p, size_t len, u8 out[32]) { Digest d = keccak256(p, len); std::memcpy(out, d.data(), 32); }
// leaf hashes of a column-major matrix (cols x rows base elements): out[i] = hash_elements(row i)
void orc_hash_rows(const u64* colmajor, size_t rows, size_t cols, u8* out) {
  std::vector<F1> row(cols);
  for (size_t i = 0; i < rows; i++) { for (size_t j = 0; j < cols; j++) row[j] = F1(colmajor[j * rows + i]); Digest d = hash_elements(row); std::memcpy(out + 32 * i, d.data(), 32); }
}
// MerkleTree::new over n 32-byte leaves; nodes_out (optional) receives the n node slots (slot 0 zero, slot 1 = root)
void orc_merkle(const u8* leaves, size_t n, u8 root[32], u8* nodes_out) {
  std::vector<Digest> lv(n); for (size_t i = 0; i < n; i++) std::memcpy(lv[i].data(), leaves + 32 * i, 32);
  MerkleTree t(std::move(lv)); std::memcpy(root, t.root().data(), 32);
  if (nodes_out) for (size_t i = 0; i < n; i++) std::memcpy(nodes_out + 32 * i, t.nodes[i].data(), 32);
}
// serialised batch proof (serialize_nodes) for `indexes`; returns length or -1
long orc_merkle_prove_batch(const u8* leaves, size_t n, const u64* indexes, size_t k, u8* out, size_t cap) {
  try {
    std::vector<Digest> lv(n); for (size_t i = 0; i < n; i++) std::memcpy(lv[i].data(), leaves + 32 * i, 32);
    MerkleTree t(std::move(lv)); std::vector<size_t> idx(indexes, indexes + k);
    std::vector<u8> b = t.prove_batch(idx).serialize_nodes();
    if (b.size() > cap) return -1;
    std::memcpy(out, b.data(), b.size()); return (long)b.size();
  } catch (...) { return -1; }
}

// src/burn_mint_prover.rs:62-107: inputs -> 12 public inputs + 4 AIR constants (txn, rcpt, nullifier, commitment) + secret element
int orc_pack_inputs(u64 burn, u64 mint, const u8 txp[32], const u8* rcpt, size_t rcpt_len, const u8* secret, size_t secret_len,
                    u32 network_id, u32 target_chain, u32 version, u64 pi_out[12], u64 consts_out[4], u64* secret_elem, char* err, size_t errcap) {
  PublicInputs pi; u64 se = 0;
  std::string e = pack_inputs(burn, mint, txp, rcpt, rcpt_len, secret, secret_len, network_id, target_chain, version, pi, se);
  if (!e.empty()) { set_err(err, errcap, e); return 1; }
  AirConsts ac = air_consts(pi, se);
  for (int i = 0; i < 12; i++) pi_out[i] = pi.v[i];
  consts_out[0] = ac.txn; consts_out[1] = ac.rcpt; consts_out[2] = ac.nullifier; consts_out[3] = ac.commitment;
  if (secret_elem) *secret_elem = se;
  return 0;
}
void orc_build_trace(const u64 pi[12], const u64 consts[4], size_t n, u64* out_colmajor) {
  auto t = build_trace(pi_from(pi), ac_from(consts), n);
  for (size_t j = 0; j < t.size(); j++) for (size_t i = 0; i < n; i++) out_colmajor[j * n + i] = t[j][i].v;
}
int orc_validate_options(const uint32_t o[6], char* err, size_t errcap) { std::string e = opts_from(o).validate(); if (!e.empty()) { set_err(err, errcap, e); return 1; } return 0; }

// full proof.  trace: column-major 7 x n canonical u64.  stage_ms: ST_COUNT doubles or NULL.  keep_debug != 0 keeps intermediates.
int orc_prove(const u64* trace, unsigned n_log2, const u64 pi[12], const u64 consts[4], const uint32_t o[6], u8* out, size_t cap, size_t* out_len,
              double* stage_ms, int keep_debug, char* err, size_t errcap) {
  try {
    ProofOptions opt = opts_from(o); std::string e = opt.validate(); if (!e.empty()) { set_err(err, errcap, e); return 1; }
    size_t n = size_t(1) << n_log2;
    std::vector<std::vector<F1>> t(XFG_TRACE_WIDTH, std::vector<F1>(n));
    for (size_t j = 0; j < XFG_TRACE_WIDTH; j++) for (size_t i = 0; i < n; i++) { if (trace[j * n + i] >= P) { set_err(err, errcap, "non-canonical trace element"); return 1; } t[j][i] = F1(trace[j * n + i]); }
    StageTimes st; std::vector<u8> bytes;
    if (opt.ext == XFG_EXT_NONE) bytes = prove<F1>(t, pi_from(pi), ac_from(consts), opt, &st, keep_debug ? &g_dbg1 : nullptr);
    else if (opt.ext == XFG_EXT_CUBIC) bytes = prove<F3>(t, pi_from(pi), ac_from(consts), opt, &st, keep_debug ? &g_dbg3 : nullptr);
    else bytes = prove<F2>(t, pi_from(pi), ac_from(consts), opt, &st, keep_debug ? &g_dbg2 : nullptr);
    if (keep_debug) g_dbg_ext = opt.ext;
    if (stage_ms) for (int i = 0; i < ST_COUNT; i++) stage_ms[i] = st.ms[i];
    *out_len = bytes.size();
    if (bytes.size() > cap) { set_err(err, errcap, "output buffer too small"); return 2; }
    std::memcpy(out, bytes.data(), bytes.size());
    return 0;
  } catch (const std::exception& ex) { set_err(err, errcap, ex.what()); return 3; }
}
const char* orc_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? STAGE_NAMES[i] : ""; }
int orc_stage_count() { return ST_COUNT; }

// intermediates of the last orc_prove(keep_debug=1): copies limbs (u64) or digest bytes; returns the number of u64 written
// (digests count 4 u64 each), or -1 if `cap` is too small / unknown name
long orc_debug_get(const char* name, u64* out, size_t cap) {
  std::string k(name); std::vector<u64> v;
  auto dg = [&](const Digest& d) { for (int i = 0; i < 4; i++) v.push_back(get_u64(d.data() + 8 * i)); };
  auto fill = [&](auto& D) {
    using DT = std::decay_t<decltype(D)>; (void)sizeof(DT);
    if (k == "trace_root") dg(D.trace_root); else if (k == "constraint_root") dg(D.constraint_root);
    else if (k == "remainder_commitment") dg(D.remainder_commitment);
    else if (k == "fri_roots") { for (auto& r : D.fri_roots) dg(r); }
    else if (k == "tcoef") to_limbs(D.tcoef, v); else if (k == "bcoef") to_limbs(D.bcoef, v); else if (k == "dcoef") to_limbs(D.dcoef, v);
    else if (k == "alphas") to_limbs(D.alphas, v); else if (k == "ood_frame") to_limbs(D.ood_frame, v); else if (k == "remainder") to_limbs(D.remainder, v);
    else if (k == "ce_evals") to_limbs(D.ce_evals, v); else if (k == "deep_evals") to_limbs(D.deep_evals, v);
    else if (k == "z") { for (int l = 0; l < std::decay_t<decltype(D.z)>::DEG; l++) v.push_back(D.z.limb(l)); }
    else if (k == "hz") { for (int l = 0; l < std::decay_t<decltype(D.hz)>::DEG; l++) v.push_back(D.hz.limb(l)); }
    else if (k == "nonce") v.push_back(D.nonce);
    else if (k == "positions") { for (size_t p : D.positions) v.push_back(p); }
    else return false;
    return true;
  };
  bool ok = (g_dbg_ext == XFG_EXT_QUADRATIC) ? fill(g_dbg2) : (g_dbg_ext == XFG_EXT_CUBIC) ? fill(g_dbg3) : fill(g_dbg1);
  if (!ok || v.size() > cap) return -1;
  std::memcpy(out, v.data(), v.size() * 8); return (long)v.size();
}

// winter-fri folding::apply_drp, folding factor F, offset 7: evals = Nl elements of `deg` limbs (element-major) -> Nl/F
void orc_fri_fold(const u64* evals, size_t Nl, int deg, size_t F, const u64* alpha, u64* out) {
  size_t rows = Nl / F; u64 gl_inv = finv(root_of_unity(ilog2(Nl))), oinv = finv(XFG_GENERATOR);
  std::vector<u64> xinv = power_series(gl_inv, rows, oinv);
  for (size_t i = 0; i < rows; i++) {
    if (deg == 1) { std::vector<F1> row(F); for (size_t j = 0; j < F; j++) row[j] = F1(evals[i + j * rows]); out[i] = fold_row<F1>(row.data(), F, xinv[i], F1(alpha[0])).v; }
    else { std::vector<F2> row(F); for (size_t j = 0; j < F; j++) row[j] = F2(evals[2 * (i + j * rows)], evals[2 * (i + j * rows) + 1]);
           F2 r = fold_row<F2>(row.data(), F, xinv[i], F2(alpha[0], alpha[1])); out[2 * i] = r.a0; out[2 * i + 1] = r.a1; }
  }
}

// returns 0 = accepted, 1 = rejected (reason in err)
int orc_verify(const u8* proof, size_t len, const u64 pi[12], const u64 consts[4], const uint32_t o[6], char* err, size_t errcap) {
  try {
    ProofOptions opt = opts_from(o); std::string e = opt.validate(); if (!e.empty()) { set_err(err, errcap, e); return 1; }
    std::string r = (opt.ext == XFG_EXT_NONE) ? verify<F1>(proof, len, pi_from(pi), ac_from(consts), opt)
                  : (opt.ext == XFG_EXT_CUBIC) ? verify<F3>(proof, len, pi_from(pi), ac_from(consts), opt)
                                               : verify<F2>(proof, len, pi_from(pi), ac_from(consts), opt);
    if (!r.empty()) { set_err(err, errcap, r); return 1; }
    return 0;
  } catch (const std::exception& ex) { set_err(err, errcap, ex.what()); return 1; }
}

// ---- generic degree-<=2 AIR ----
// desc: width, num_pub, num_const, num_instr, num_out, num_assert; code = 3 u32 per instruction (op, a, b); assertions = 3 u64 each (column, step, value)
static AirDef air_from(const uint32_t desc[6], const u64* pub, const u64* consts, const uint32_t* code, const uint32_t* outs, const u64* asr) {
  AirDef a; a.width = desc[0]; a.pub_inputs.assign(pub, pub + desc[1]); a.constants.assign(consts, consts + desc[2]);
  for (uint32_t i = 0; i < desc[3]; i++) a.code.push_back(Instr{code[3 * i], code[3 * i + 1], code[3 * i + 2]});
  a.outputs.assign(outs, outs + desc[4]);
  for (uint32_t i = 0; i < desc[5]; i++) a.assertions.push_back(Assertion{(u32)asr[3 * i], (size_t)asr[3 * i + 1], asr[3 * i + 2]});
  return a;
}
int orc_prove_air(const u64* trace, unsigned n_log2, const uint32_t desc[6], const u64* pub, const u64* consts, const uint32_t* code, const uint32_t* outs,
                  const u64* asr, const uint32_t o[6], u8* out, size_t cap, size_t* out_len, double* stage_ms, int keep_debug, char* err, size_t errcap) {
  try {
    ProofOptions opt = opts_from(o); std::string e = opt.validate(); if (!e.empty()) { set_err(err, errcap, e); return 1; }
    const size_t n = size_t(1) << n_log2;
    AirDef air = air_from(desc, pub, consts, code, outs, asr);
    e = validate_air(air, n); if (!e.empty()) { set_err(err, errcap, e); return 1; }
    std::vector<std::vector<F1>> t(air.width, std::vector<F1>(n));
    for (size_t j = 0; j < air.width; j++) for (size_t i = 0; i < n; i++) { if (trace[j * n + i] >= P) { set_err(err, errcap, "non-canonical trace element"); return 1; } t[j][i] = F1(trace[j * n + i]); }
    StageTimes st; std::vector<u8> bytes;
    if (opt.ext == XFG_EXT_NONE) bytes = prove<F1>(t, air, opt, &st, keep_debug ? &g_dbg1 : nullptr);
    else if (opt.ext == XFG_EXT_CUBIC) bytes = prove<F3>(t, air, opt, &st, keep_debug ? &g_dbg3 : nullptr);
    else bytes = prove<F2>(t, air, opt, &st, keep_debug ? &g_dbg2 : nullptr);
    if (keep_debug) g_dbg_ext = opt.ext;
    if (stage_ms) for (int i = 0; i < ST_COUNT; i++) stage_ms[i] = st.ms[i];
    *out_len = bytes.size();
    if (bytes.size() > cap) { set_err(err, errcap, "output buffer too small"); return 2; }
    std::memcpy(out, bytes.data(), bytes.size());
    return 0;
  } catch (const std::exception& ex) { set_err(err, errcap, ex.what()); return 3; }
}
int orc_verify_air(const u8* proof, size_t len, const uint32_t desc[6], const u64* pub, const u64* consts, const uint32_t* code, const uint32_t* outs,
                   const u64* asr, const uint32_t o[6], char* err, size_t errcap) {
  try {
    ProofOptions opt = opts_from(o); std::string e = opt.validate(); if (!e.empty()) { set_err(err, errcap, e); return 1; }
    AirDef air = air_from(desc, pub, consts, code, outs, asr);
    std::string r = (opt.ext == XFG_EXT_NONE) ? verify<F1>(proof, len, air, opt) : (opt.ext == XFG_EXT_CUBIC) ? verify<F3>(proof, len, air, opt) : verify<F2>(proof, len, air, opt);
    if (!r.empty()) { set_err(err, errcap, r); return 1; }
    return 0;
  } catch (const std::exception& ex) { set_err(err, errcap, ex.what()); return 1; }
}

}  // extern "C"
