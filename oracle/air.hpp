// oracle/air.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// The *normalised* BurnMintAir: the reference's AIR with its defects 1-6 removed, and
// nothing else changed.  Follows the reference line by line:
//   public inputs + ToElements order ........ src/burn_mint_air.rs:23-71
//   Keccak scalars (nullifier, commitment) .. src/burn_mint_air.rs:124-133, 157-202
//   transition constraints .................. src/burn_mint_air.rs:205-268, 335-378
//   assertions .............................. src/burn_mint_air.rs:380-395 (last step n-1 instead of the literal 63)
//   trace ................................... src/burn_mint_air.rs:442-476 (state = floor(4i/n))
//   input validation / packing .............. src/burn_mint_prover.rs:74-107, 132-221
//   ProofOptions + Context (winter-air 0.8.3) SURVEY.md A.2, A.4, A.12
#pragma once
#include <algorithm>
#include <string>
#include "hash.hpp"

namespace orc {

struct ProofOptions {
  u32 num_queries = XFG_DEF_NUM_QUERIES, blowup = XFG_DEF_BLOWUP, grinding = XFG_DEF_GRINDING, ext = XFG_EXT_NONE,
      folding = XFG_DEF_FRI_FOLDING, rem_max_deg = XFG_DEF_FRI_REM_MAX;
  // ProofOptions::new range checks (A.2, D); returns an error string or ""
  std::string validate() const {
    auto pow2 = [](u32 x) { return x && !(x & (x - 1)); };
    if (num_queries < 1) return "number of queries must be greater than 0";
    if (num_queries > 255) return "number of queries cannot be greater than 255";
    if (!pow2(blowup)) return "blowup factor must be a power of 2";
    if (blowup < 2) return "blowup factor cannot be smaller than 2";
    if (blowup > 128) return "blowup factor cannot be greater than 128";
    if (grinding > 32) return "grinding factor cannot be greater than 32";
    if (!pow2(folding)) return "FRI folding factor must be a power of 2";
    if (folding < 2 || folding > 16) return "FRI folding factor out of range";
    if (rem_max_deg > 255 || !pow2(rem_max_deg + 1)) return "FRI polynomial remainder degree must be one less than a power of two";
    if (ext != XFG_EXT_NONE && ext != XFG_EXT_QUADRATIC && ext != XFG_EXT_CUBIC) return "UnsupportedFieldExtension";
    return "";
  }
  // ProofOptions::to_elements (A.2, D)
  std::vector<F1> to_elements() const {
    return {F1((u64)ext << 16 | (u64)folding << 8 | rem_max_deg), F1(grinding), F1(blowup), F1(num_queries)};
  }
  // write_into: 6 bytes (A.2, D)
  void write_into(std::vector<u8>& o) const { for (u32 b : {num_queries, blowup, grinding, ext, folding, rem_max_deg}) o.push_back((u8)b); }
  // FriOptions::num_fri_layers (A.10, D)
  size_t num_fri_layers(size_t domain) const {
    size_t r = 0, mx = (size_t)(rem_max_deg + 1) * blowup;
    while (domain > mx) { domain /= folding; r++; }
    return r;
  }
};

struct PublicInputs { u64 v[XFG_NUM_PUB_INPUTS]; };   // order of src/burn_mint_air.rs:54-71
struct AirConsts { u64 txn, rcpt, nullifier, commitment; };

inline void le64(std::vector<u8>& b, u64 v) { put_u64(b, v); }
inline void ascii(std::vector<u8>& b, const char* s) { while (*s) b.push_back((u8)*s++); }
inline u32 head_u32(const Digest& d) { return (u32)d[0] | (u32)d[1] << 8 | (u32)d[2] << 16 | (u32)d[3] << 24; }

// src/burn_mint_prover.rs:211-221
inline u32 prover_recipient_hash(const u8* addr, size_t len) { std::vector<u8> b(addr, addr + len); ascii(b, "recipient"); return head_u32(keccak256(b)); }
// src/burn_mint_air.rs:124-133
inline u64 compute_nullifier(const PublicInputs& pi, u64 secret) {
  std::vector<u8> b; le64(b, secret); ascii(b, "nullifier"); le64(b, pi.v[XFG_PI_BURN]); return head_u32(keccak256(b));
}
// src/burn_mint_air.rs:157-170
inline Digest air_recipient_hash(const PublicInputs& pi) {
  std::vector<u8> b; le64(b, pi.v[XFG_PI_RECIPIENT_HASH]); ascii(b, "ethereum-recipient"); ascii(b, "fuego-to-heat-bridge"); return keccak256(b);
}
// src/burn_mint_air.rs:174-202
inline u64 compute_commitment(const PublicInputs& pi, u64 secret) {
  std::vector<u8> b; le64(b, secret); le64(b, pi.v[XFG_PI_BURN]); le64(b, pi.v[XFG_PI_MINT]);
  le64(b, pi.v[XFG_PI_TXP0]); le64(b, pi.v[XFG_PI_TXP1]); le64(b, pi.v[XFG_PI_TXP2]); le64(b, pi.v[XFG_PI_TXP3]);
  Digest r = air_recipient_hash(pi); b.insert(b.end(), r.begin(), r.end());
  le64(b, pi.v[XFG_PI_NETWORK_ID]); le64(b, pi.v[XFG_PI_TARGET_CHAIN]); le64(b, pi.v[XFG_PI_VERSION]);
  ascii(b, "heat-commitment-v1"); return head_u32(keccak256(b));
}
inline AirConsts air_consts(const PublicInputs& pi, u64 secret) {
  return {(u64)(u32)pi.v[XFG_PI_TXN_HASH], (u64)(u32)pi.v[XFG_PI_RECIPIENT_HASH], compute_nullifier(pi, secret), compute_commitment(pi, secret)};
}

// src/burn_mint_prover.rs:62-107 + 132-208.  Returns "" or the reference's error text.
inline std::string pack_inputs(u64 burn, u64 mint, const u8 txp[32], const u8* rcpt, size_t rcpt_len, const u8* secret, size_t secret_len,
                               u32 network_id, u32 target_chain, u32 version, PublicInputs& pi, u64& secret_elem) {
  u64 legacy = get_u64(txp);
  if (burn != XFG_STD_BURN && burn != XFG_LARGE_BURN) return "Burn amount must be exactly 0.8 XFG (8,000,000 atomic units) or 800 XFG (8,000,000,000 atomic units)";
  if (mint != burn) return "Mint amount does not match burn amount for 1:1 atomic unit conversion";
  if (legacy == 0) return "Transaction hash must be greater than 0";
  if (rcpt_len != 20) return "Recipient address must be exactly 20 bytes";
  if (secret_len < 4) return "Secret must be at least 4 bytes";
  if (secret_len < 8) return "Secret must be at least 8 bytes";   // the reference panics here (:201, defect B.1-9)
  secret_elem = (u32)get_u64(secret);
  auto w = [&](int o) { return (u64)((u32)txp[o] | (u32)txp[o + 1] << 8 | (u32)txp[o + 2] << 16 | (u32)txp[o + 3] << 24); };
  pi.v[XFG_PI_BURN] = (u32)burn; pi.v[XFG_PI_MINT] = (u32)mint; pi.v[XFG_PI_TXN_HASH] = (u32)legacy;
  pi.v[XFG_PI_RECIPIENT_HASH] = prover_recipient_hash(rcpt, rcpt_len); pi.v[XFG_PI_STATE] = 0;
  pi.v[XFG_PI_TXP0] = w(0); pi.v[XFG_PI_TXP1] = w(4); pi.v[XFG_PI_TXP2] = w(8); pi.v[XFG_PI_TXP3] = w(12);
  pi.v[XFG_PI_NETWORK_ID] = network_id; pi.v[XFG_PI_TARGET_CHAIN] = target_chain; pi.v[XFG_PI_VERSION] = version;
  return "";
}

// src/burn_mint_air.rs:442-476, generalised to n rows (B.2): column-major, 7 columns
inline std::vector<std::vector<F1>> build_trace(const PublicInputs& pi, const AirConsts& c, size_t n) {
  std::vector<std::vector<F1>> t(XFG_TRACE_WIDTH, std::vector<F1>(n));
  for (size_t i = 0; i < n; i++) {
    t[0][i] = F1(pi.v[XFG_PI_BURN]); t[1][i] = F1(pi.v[XFG_PI_MINT]); t[2][i] = F1(pi.v[XFG_PI_TXN_HASH]);
    t[3][i] = F1(pi.v[XFG_PI_RECIPIENT_HASH]); t[4][i] = F1((u64)(4 * i / n)); t[5][i] = F1(c.nullifier); t[6][i] = F1(c.commitment);
  }
  return t;
}

// src/burn_mint_air.rs:335-378; E = F1 on the prover's constraint domain, E = extension at the OOD point
template <class E> inline void evaluate_transition(const E* cur, const E* nxt, const AirConsts& c, E* r) {
  E std_burn = E::from_base(XFG_STD_BURN), large = E::from_base(XFG_STD_BURN) * E::from_base(1000);
  r[0] = (cur[0] - std_burn) * (cur[0] - large);        // :207-219
  r[1] = cur[1] - cur[0];                               // :231
  r[2] = cur[2] - E::from_base(c.txn);                  // :362
  r[3] = cur[3] - E::from_base(c.rcpt);                 // :365
  E d = nxt[4] - cur[4]; r[4] = d * (d - E::one());     // :240-246
  r[5] = cur[5] - E::from_base(c.nullifier);            // :264-267
  r[6] = cur[6] - E::from_base(c.commitment);           // :376-377
}

// src/burn_mint_air.rs:380-395 in Winterfell's sorted order (stride, first_step, column) (A.8, D)
struct Assertion { u32 column; size_t step; u64 value; };
inline std::vector<Assertion> get_assertions(const PublicInputs& pi, const AirConsts& c, size_t n) {
  return {{0, 0, pi.v[XFG_PI_BURN]}, {1, 0, pi.v[XFG_PI_MINT]}, {2, 0, pi.v[XFG_PI_TXN_HASH]}, {3, 0, pi.v[XFG_PI_RECIPIENT_HASH]},
          {4, 0, 0}, {5, 0, c.nullifier}, {6, 0, c.commitment}, {4, n - 1, XFG_FINAL_STATE}};
}

// ---- generic AIR (transition degrees up to 9 = up to 8 composition columns): a straight-line program over the evaluation frame --------------------
// Value ids: [0, w) = current row, [w, 2w) = next row, [2w, 2w + C) = constants, 2w + C + i = result of instruction i.
// This is the role of a user's `Air::evaluate_transition` body (e.g. the 4-column XfgBurnAir sketch, src/winterfell_air.rs:87-127);
// Winterfell's own rules around it - assertion ordering (A.8), ce_blowup 2 and one composition column for degrees <= 2 (A.3) -
// are unchanged.
enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2 };
struct Instr { u32 op, a, b; };
struct AirDef {
  bool burn_mint = false;                   // the hard-wired normalised BurnMintAir (AirConsts in `ac`), else the program below
  AirConsts ac{};
  size_t width = 0, num_transition = 0;
  std::vector<u64> pub_inputs;              // ToElements order; appended to the coin seed after Context::to_elements
  std::vector<u64> constants; std::vector<Instr> code; std::vector<u32> outputs;   // outputs[j] = value id of constraint j
  std::vector<Assertion> assertions;        // Winterfell's sorted order (stride, first_step, column) = (step, column) for single assertions
  u32 max_degree = 2;                       // highest transition-constraint degree d (computed by validate_air); winter-air derives from it (A.3):
  // ce_blowup = max(2, next_pow2(d - 1))   (TransitionConstraintDegree::min_blowup_factor; pinned by running the reference binary with declared degrees 3, 4, 5)
  size_t ce_blowup() const { size_t c = 2; while (c + 1 < max_degree) c <<= 1; return c; }
  // composition columns = max(1, d - 1)    (AirContext::num_constraint_composition_columns: ceil((d (n-1) - (n-1)) / n) for n > d)
  size_t comp_columns() const { return max_degree > 2 ? max_degree - 1 : 1; }
};
inline AirDef burn_mint_air(const PublicInputs& pi, const AirConsts& c, size_t n) {
  AirDef a; a.burn_mint = true; a.ac = c; a.width = XFG_TRACE_WIDTH; a.num_transition = XFG_NUM_TRANSITION;
  a.pub_inputs.assign(pi.v, pi.v + XFG_NUM_PUB_INPUTS); a.assertions = get_assertions(pi, c, n); return a;
}
// Air::new-style validation of a generic definition; returns "" or the reason (strings follow winter-air's panics where one exists)
inline std::string validate_air(AirDef& a, size_t n) {
  if (a.burn_mint) return "";
  const size_t w = a.width, C = a.constants.size();
  if (w < 1 || w > 255) return "number of columns must be between 1 and 255";
  if (a.outputs.empty()) return "at least one transition constraint degree must be specified";
  if (a.assertions.empty()) return "at least one assertion must be specified";
  for (u64 c : a.constants) if (c >= P) return "non-canonical constant";
  for (u64 c : a.pub_inputs) if (c >= P) return "non-canonical public input";
  std::vector<u32> deg(2 * w + C + a.code.size(), 0);
  for (size_t i = 0; i < 2 * w; i++) deg[i] = 1;
  for (size_t i = 0; i < a.code.size(); i++) {
    const Instr& in = a.code[i]; const size_t id = 2 * w + C + i;
    if (in.a >= id || in.b >= id || in.op > OP_MUL) return "invalid instruction";
    deg[id] = in.op == OP_MUL ? deg[in.a] + deg[in.b] : std::max(deg[in.a], deg[in.b]);
    if (deg[id] > XFG_AIR_MAX_DEGREE) return "transition constraint degree above 9 is not supported";
  }
  a.max_degree = 1;
  for (u32 o : a.outputs) { if (o >= deg.size()) return "invalid constraint output"; if (deg[o] == 0) return "transition constraint degree must be at least one"; a.max_degree = std::max(a.max_degree, deg[o]); }
  if (a.max_degree >= n) return "transition constraint degree must be smaller than the trace length";
  std::sort(a.assertions.begin(), a.assertions.end(), [](const Assertion& x, const Assertion& y) { return x.step != y.step ? x.step < y.step : x.column < y.column; });
  for (size_t i = 0; i < a.assertions.size(); i++) {
    const Assertion& s = a.assertions[i];
    if (s.column >= w) return "assertion column out of range";
    if (s.step >= n) return "assertion step out of range";
    if (s.value >= P) return "non-canonical assertion value";
    if (i && a.assertions[i - 1].step == s.step && a.assertions[i - 1].column == s.column) return "duplicate assertion";
  }
  a.num_transition = a.outputs.size();
  return "";
}
template <class E> inline void eval_air_transition(const AirDef& a, const E* cur, const E* nxt, E* r) {
  if (a.burn_mint) { evaluate_transition<E>(cur, nxt, a.ac, r); return; }
  const size_t w = a.width, C = a.constants.size();
  std::vector<E> v(2 * w + C + a.code.size());
  for (size_t i = 0; i < w; i++) { v[i] = cur[i]; v[w + i] = nxt[i]; }
  for (size_t i = 0; i < C; i++) v[2 * w + i] = E::from_base(a.constants[i]);
  for (size_t i = 0; i < a.code.size(); i++) {
    const Instr& in = a.code[i];
    v[2 * w + C + i] = in.op == OP_ADD ? v[in.a] + v[in.b] : in.op == OP_SUB ? v[in.a] - v[in.b] : v[in.a] * v[in.b];
  }
  for (size_t j = 0; j < a.outputs.size(); j++) r[j] = v[a.outputs[j]];
}

// winter-air Context::to_elements followed by the public inputs = coin seed elements (A.4, D)
inline std::vector<F1> seed_elements(size_t n, const ProofOptions& o, size_t width, const std::vector<u64>& pub_inputs) {
  std::vector<F1> e;
  e.push_back(F1((u64)width << 8));                    // (main_width << 8) | num_aux_segments
  e.push_back(F1(P & 0xFFFFFFFFULL)); e.push_back(F1(P >> 32));   // modulus LE bytes, two halves
  for (auto x : o.to_elements()) e.push_back(x);
  e.push_back(F1((u64)(u32)n));
  for (u64 v : pub_inputs) e.push_back(F1(v));
  return e;
}
// winter-air Context::write_into (A.12, D)
inline void write_context(std::vector<u8>& o, size_t n, const ProofOptions& opt, size_t width) {
  unsigned lg = 0; while ((size_t(1) << lg) < n) lg++;
  o.push_back((u8)width); o.push_back(0); o.push_back(0); o.push_back((u8)lg);
  o.push_back(0); o.push_back(0);        // u16 meta_len = 0
  o.push_back(8); put_u64(o, P);          // modulus
  opt.write_into(o);
}

}  // namespace orc
