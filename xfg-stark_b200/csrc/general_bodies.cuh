// general_bodies.cuh — the general-options proof pipeline: every `ProofOptions` value the reference's `XfgBurnMintProver::with_options`
// (src/burn_mint_prover.rs:44-49) accepts that the tuned 8/8 pipeline (prover.cu) does not serve: blowup factors 2 .. 128, FRI folding
// factors 2 / 4 / 8 / 16, any remainder degree, and `FieldExtension::Cubic`.
//
// Replaces the same winter-prover 0.8.3 `Prover::prove` stages as prover.cu (A.4-A.12), for an AIR given as a compiled
// program (generic_air.cuh; the burn-mint AIR is one such program).  Design: one GPU thread per output value, no cooperation inside a
// block - every kernel is a functor `body(t)` over a flat index space, launched through `go_kernel` (general.cu).  The same bodies compile
// as plain C++ (all field / BLAKE3 helpers are host+device), which is how tests/host_emul runs this pipeline on a CPU-only box against
// the oracle and the reference's own proofs before any GPU time is spent; the product library only ever launches them as CUDA kernels.
// The heavy transforms (interpolation, LDE) and the Merkle levels above the leaves are the tuned kernels of ntt.cu / merkle.cu.
//
// Data layout: LDE matrices are coset-major like the tuned pipeline, [column][k][m] with LDE row i = B m + k (k < B = blowup); the DEEP /
// FRI evaluations are natural order [limb][i]; Merkle trees are heap-ordered (tree[M + i] = leaf i, root tree[1]).
#pragma once
#include "generic_air.cuh"
#include "ntt.cuh"

namespace xfg {

static constexpr int GO_MAX_LAYERS = 32;      // folding factor 2 on a 2^27-point domain
static constexpr int GO_OOD_CHUNKS = 2048;    // partial sums per polynomial of the out-of-domain evaluation (fewer for traces shorter than that)
static constexpr int GO_MAX_EXT = 3;
static constexpr int GO_MAX_CE = 8;           // constraint-evaluation blowup (and composition columns) for transition degrees up to XFG_AIR_MAX_DEGREE = 9

#if defined(__CUDACC__)
#define GO_NOINLINE __host__ __device__ __noinline__
#else
#define GO_NOINLINE inline
#endif

// device-resident state of a general-options proof (the counterpart of ProofState + GenState, with room for cubic elements)
struct GoState {
  // init block: one host->device copy per proof (coin seed elements, cleared flags, unset nonce)
  u64 seed_limbs[MAX_SEED_LIMBS]; u32 seed_count; u32 error_flags; unsigned long long nonce;
  // coin
  Digest seed; u64 counter;
  Digest trace_root, constraint_root, fri_roots[GO_MAX_LAYERS], remainder_commitment;
  u64 coef[GEN_MAX_CONSTRAINTS + GEN_MAX_ASSERTIONS][GO_MAX_EXT];   // transition coefficients, then boundary coefficients (A.8)
  u64 z[GO_MAX_EXT], zg[GO_MAX_EXT], hz[GO_MAX_CE][GO_MAX_EXT];     // hz[i] = H_i(z), one per composition column
  u64 ood_frame[2 * GEN_MAX_WIDTH][GO_MAX_EXT];                     // T_0(z), T_0(zg), T_1(z), ... (A.9)
  u64 dcoef[GEN_MAX_WIDTH + GO_MAX_CE][GO_MAX_EXT];                 // trace columns, then composition columns
  u64 deep_c1[GO_MAX_EXT], deep_c2[GO_MAX_EXT];
  u64 alphas[GO_MAX_LAYERS][GO_MAX_EXT];
  u64 remainder[MAX_REMAINDER][GO_MAX_EXT]; u32 remainder_len;
  u32 num_positions; u32 positions[MAX_Q];
  u32 fri_num_positions[GO_MAX_LAYERS]; u32 fri_positions[GO_MAX_LAYERS][MAX_Q];
};
static constexpr size_t GO_INIT_BYTES = sizeof(u64) * MAX_SEED_LIMBS + 16;
static_assert(GO_INIT_BYTES == offsetof(GoState, seed), "init block layout");

// ------------------------------------------------------------------------------------------------------------------
// BLAKE3 (winter-crypto Blake3_256, A.6) on ONE shared, non-inlined compression: these bodies are not the hashing hot path
// ------------------------------------------------------------------------------------------------------------------
static GO_NOINLINE void go_compress(const u32* cv, const u32* m, u32 len, u32 flags, u32 ctr, u32* out) { b3_compress(cv, m, len, flags, out, ctr); }
XFG_HD Digest go_merge(const Digest& l, const Digest& r) {
  u32 cv[8]; b3_iv(cv); u32 m[16];
  for (int i = 0; i < 8; i++) { m[i] = l.w[i]; m[8 + i] = r.w[i]; }
  Digest d; go_compress(cv, m, 64, B3_SINGLE, 0, d.w); return d;
}
XFG_HD Digest go_merge_int(const Digest& s, u64 v) {
  u32 cv[8]; b3_iv(cv); u32 m[16];
  for (int i = 0; i < 8; i++) { m[i] = s.w[i]; m[8 + i] = 0; }
  m[8] = (u32)v; m[9] = (u32)(v >> 32);
  Digest d; go_compress(cv, m, 40, B3_SINGLE, 0, d.w); return d;
}
XFG_HD void go_parent(const u32* l, const u32* r, bool root, u32* out) {
  u32 cv[8]; b3_iv(cv); u32 m[16];
  for (int i = 0; i < 8; i++) { m[i] = l[i]; m[8 + i] = r[i]; }
  go_compress(cv, m, 64, XFG_B3_PARENT | (root ? XFG_B3_ROOT : 0), 0, out);
}
// hash_elements of nl limbs (limb i = get(i), 8 bytes LE each), any length: chunks of 128 limbs, BLAKE3 tree mode with the standard
// chaining-value stack (a completed subtree of 2^k chunks is merged as soon as its sibling arrives)
template <class Get> XFG_HD Digest go_hash_stream(int nl, Get get) {
  const int chunks = nl <= 128 ? 1 : (nl + 127) / 128;
  u32 stack[12][8]; int sp = 0;
  Digest out;
  for (int c = 0; c < chunks; c++) {
    int cl = nl - c * 128; if (cl > 128) cl = 128;
    const int nb = cl <= 0 ? 1 : (cl + 7) / 8;
    u32 cv[8]; b3_iv(cv);
    for (int b = 0; b < nb; b++) {
      u32 m[16];
      for (int i = 0; i < 8; i++) { const int li = b * 8 + i; const u64 v = li < cl ? get(c * 128 + li) : 0; m[2 * i] = (u32)v; m[2 * i + 1] = (u32)(v >> 32); }
      const int rem = cl - b * 8; const u32 len = rem >= 8 ? 64 : (rem > 0 ? (u32)rem * 8 : 0);
      const u32 flags = (b == 0 ? XFG_B3_CHUNK_START : 0) | (b == nb - 1 ? (XFG_B3_CHUNK_END | (chunks == 1 ? XFG_B3_ROOT : 0)) : 0);
      go_compress(cv, m, len, flags, (u32)c, cv);
    }
    if (c + 1 < chunks) {
      u32 total = (u32)c + 1;
      while ((total & 1u) == 0) { sp--; go_parent(stack[sp], cv, false, cv); total >>= 1; }
      for (int i = 0; i < 8; i++) stack[sp][i] = cv[i];
      sp++;
    } else {
      while (sp > 0) { sp--; go_parent(stack[sp], cv, sp == 0, cv); }
      for (int i = 0; i < 8; i++) out.w[i] = cv[i];
    }
  }
  return out;
}

// ------------------------------------------------------------------------------------------------------------------
// coin (winter-crypto DefaultRandomCoin, A.5), serial: the transcript steps are single-thread bodies
// ------------------------------------------------------------------------------------------------------------------
struct GoCoin { Digest seed; u64 counter; };
XFG_HD GoCoin go_coin_load(const GoState* s) { GoCoin c; c.seed = s->seed; c.counter = s->counter; return c; }
XFG_HD void go_coin_store(GoState* s, const GoCoin& c) { s->seed = c.seed; s->counter = c.counter; }
XFG_HD void go_reseed(GoCoin& c, const Digest& d) { c.seed = go_merge(c.seed, d); c.counter = 0; }
// draw::<E>(): the first 8 D bytes of next(); a candidate with a non-canonical limb is skipped (at most 1000 tries)
template <int D> XFG_HD bool go_draw(GoCoin& c, u64* out) {
  for (int t = 0; t < XFG_COIN_MAX_DRAWS; t++) {
    c.counter += 1;
    const Digest d = go_merge_int(c.seed, c.counter);
    bool ok = true;
    for (int l = 0; l < D; l++) { const u64 v = (u64)d.w[2 * l] | ((u64)d.w[2 * l + 1] << 32); if (v >= GL_P) ok = false; out[l] = v; }
    if (ok) return true;
  }
  return false;
}
XFG_HD void go_flag(u32* flags, u32 bit) {
#if defined(__CUDA_ARCH__)
  atomicOr(flags, bit);
#else
  *flags |= bit;
#endif
}
template <int D> XFG_HD Ext<D> go_ld(const u64* p) { Ext<D> r; for (int l = 0; l < D; l++) r.set_limb(l, p[l]); return r; }
template <int D> XFG_HD void go_st(u64* p, const Ext<D>& v) { for (int l = 0; l < GO_MAX_EXT; l++) p[l] = l < D ? v.limb(l) : 0; }
template <int D> XFG_HD Ext<D> go_pow(Ext<D> b, u64 e) { Ext<D> r(1); while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; } return r; }

// ------------------------------------------------------------------------------------------------------------------
// leaves: hash_elements of LDE row i = B m + k of a coset-major matrix of `limbs` arrays (A.7); t = k n + m
// ------------------------------------------------------------------------------------------------------------------
struct GoLeaf {
  const u64* data; u64 limb_stride; u32 limbs, lb, ln; Digest* tree;
  XFG_HD void operator()(size_t t) const {
    const size_t n = size_t(1) << ln, k = t >> ln, m = t & (n - 1), i = (m << lb) | k;
    const u64* p = data + k * n + m; const u64 ls = limb_stride;
    tree[(n << lb) + i] = go_hash_stream((int)limbs, [p, ls](int l) { return p[(size_t)l * ls]; });
  }
};
// leaves of a FRI layer: row r = the F evaluations at r + j R (R rows), limbs interleaved per element (A.10 transpose_slice + hash_elements)
template <int D> struct GoFriLeaf {
  const u64* ev; u64 limb_stride; u32 F; u64 R; Digest* tree;
  XFG_HD void operator()(size_t r) const {
    const u64* e = ev; const u64 ls = limb_stride, rows = R;
    tree[R + r] = go_hash_stream((int)(F * D), [e, ls, rows, r](int li) { return e[(size_t)(li % D) * ls + r + (size_t)(li / D) * rows]; });
  }
};

// ------------------------------------------------------------------------------------------------------------------
// transcript steps (count = 1)
// ------------------------------------------------------------------------------------------------------------------
// coin seed = hash_elements(context || public inputs) (A.4); commit_trace; transition then boundary coefficients (A.8)
template <int D> struct GoStepTrace {
  GoState* s; const Digest* tree; u32 ncoef;
  XFG_HD void operator()(size_t) const {
    const u64* sl = s->seed_limbs;
    GoCoin c; c.seed = go_hash_stream((int)s->seed_count, [sl](int i) { return sl[i]; }); c.counter = 0;
    s->trace_root = tree[1];
    go_reseed(c, s->trace_root);
    bool ok = true;
    for (u32 i = 0; i < ncoef; i++) { u64 v[GO_MAX_EXT] = {0, 0, 0}; ok &= go_draw<D>(c, v); for (int l = 0; l < GO_MAX_EXT; l++) s->coef[i][l] = l < D ? v[l] : 0; }
    if (!ok) s->error_flags |= ERR_FLAG_COIN;
    go_coin_store(s, c);
  }
};
// commit_constraints; z; z g
template <int D> struct GoStepComp {
  GoState* s; const Digest* tree; u64 g_n;
  XFG_HD void operator()(size_t) const {
    GoCoin c = go_coin_load(s);
    s->constraint_root = tree[1];
    go_reseed(c, s->constraint_root);
    u64 v[GO_MAX_EXT] = {0, 0, 0};
    if (!go_draw<D>(c, v)) s->error_flags |= ERR_FLAG_COIN;
    const Ext<D> z = go_ld<D>(v);
    go_st<D>(s->z, z); go_st<D>(s->zg, mul_base(z, g_n));
    go_coin_store(s, c);
  }
};
// commit_fri_layer; alpha
template <int D> struct GoStepFri {
  GoState* s; const Digest* tree; u32 layer;
  XFG_HD void operator()(size_t) const {
    GoCoin c = go_coin_load(s);
    s->fri_roots[layer] = tree[1];
    go_reseed(c, s->fri_roots[layer]);
    u64 v[GO_MAX_EXT] = {0, 0, 0};
    if (!go_draw<D>(c, v)) s->error_flags |= ERR_FLAG_COIN;
    for (int l = 0; l < GO_MAX_EXT; l++) s->alphas[layer][l] = v[l];
    go_coin_store(s, c);
  }
};

// ------------------------------------------------------------------------------------------------------------------
// evaluate_constraints (A.8): over the constraint-evaluation domain = LDE cosets k' B / ce, k' < ce
//   H(x) = T(x) (x - g^(n-1)) / (x^n - 1) + sum_groups B_g(x) / (x - g^step_g)
// out: [limb][k'][m], k' < ce = the constraint-evaluation blowup (2 for degrees <= 3, 4 for 4-5, 8 for 6-9)
// ------------------------------------------------------------------------------------------------------------------
static constexpr int GO_PTS = 4;   // points per thread of the constraint and DEEP bodies: their base-field inversions are batched (Montgomery's trick)
// in-place batch inversion of cnt non-zero base-field values (one gl_inv)
XFG_HD void go_batch_inv(u64* v, int cnt) {
  u64 pre[2 * GO_PTS]; u64 acc = 1;
  for (int i = 0; i < cnt; i++) { pre[i] = acc; acc = gl_mul(acc, v[i]); }
  acc = gl_inv(acc);
  for (int i = cnt - 1; i >= 0; i--) { const u64 t = gl_mul(pre[i], acc); acc = gl_mul(acc, v[i]); v[i] = t; }
}
// t = k' (n / pts) + q: the thread evaluates points m = q + j n / pts, j < pts (pts = GO_PTS, or 1 for the shortest traces), of coset k'
template <int D> struct GoConstraint {
  const u64* lde; u32 ln, lb, lce, pts; const GenProgram* prog; const GoState* s; PowTable wn; u64 s_ce[GO_MAX_CE], zinv[GO_MAX_CE], g_last; u64* out;
  XFG_HD void operator()(size_t t) const {
    const size_t n = size_t(1) << ln, N = n << lb, per = n / pts, kp = t / per, q = t % per;
    const size_t k = kp << (lb - lce);       // constraint-evaluation coset k' = LDE coset k' B / ce
    const u64* base = lde + k * n;
    const u32 T = prog->num_constraints, A = prog->num_assertions, G = prog->num_groups, NI = prog->num_instr;
    Ext<D> u[GO_PTS], num[GO_PTS]; u64 den[GO_PTS];
    for (u32 j = 0; j < pts; j++) {
      const size_t m = q + j * per, mn = (m + 1) & (n - 1);
      u64 slot[GEN_MAX_SLOTS];
      Ext<D> ts;
      for (u32 i = 0; i < NI; i++) {
        const GenInstr in = prog->code[i];
        const u32 op = in.w0 & 15u, dst = in.w0 >> 8;
        u64 v[2];
        for (int o = 0; o < 2; o++) {
          const u32 kind = (in.w0 >> (4 + 2 * o)) & 3u, idx = o ? in.w1 >> 16 : in.w1 & 0xFFFFu;
          v[o] = kind == GK_SLOT ? slot[idx] : kind == GK_CONST ? prog->constants[idx] : base[(size_t)idx * N + (kind == GK_CUR ? m : mn)];
          if (op == GOP_OUT) break;
        }
        if (op == GOP_OUT) { ts = ts + mul_base(go_ld<D>(s->coef[dst]), v[0]); continue; }
        slot[dst] = op == GOP_MUL ? gl_mul(v[0], v[1]) : op == GOP_ADD ? gl_add(v[0], v[1]) : gl_sub(v[0], v[1]);
      }
      const u64 x = gl_mul(s_ce[kp], pow_lookup(wn, m));
      Ext<D> nm; u64 dn = 1; u32 ai = 0;
      for (u32 g = 0; g < G; g++) {      // the boundary sum as one fraction: one inversion per point whatever the number of divisors
        Ext<D> bs;
        for (; ai < A && prog->asr[ai].group == g; ai++)
          bs = bs + mul_base(go_ld<D>(s->coef[T + ai]), gl_sub(base[(size_t)prog->asr[ai].column * N + m], prog->asr[ai].value));
        const u64 xg = gl_sub(x, prog->group_point[g]);
        nm = mul_base(nm, xg) + mul_base(bs, dn);
        dn = gl_mul(dn, xg);
      }
      u[j] = mul_base(ts, gl_mul(gl_sub(x, g_last), zinv[kp])); num[j] = nm; den[j] = dn;
    }
    go_batch_inv(den, (int)pts);
    for (u32 j = 0; j < pts; j++) {
      const Ext<D> h = u[j] + mul_base(num[j], den[j]);
      for (int l = 0; l < D; l++) out[(((size_t)l << lce) + kp) * n + q + j * per] = h.limb(l);
    }
  }
};
// composition coefficients from the ce un-scaled coset interpolants A_c (c < ce): on coset c, x^n = 7^n w_ce^c, so for the composition polynomial
// H = sum_i x^(i n) H_i (H_i of degree < n):  A_c = sum_i (7^n w_ce^c)^i H_i,  i.e.  H_i[j] = 7^(-n i) / ce * sum_c w_ce^(-c i) A_c[j]  (a size-ce inverse
// DFT per coefficient).  Columns i < K are the composition columns (CompositionPoly::new: coefficients [i n, (i + 1) n)); the others must vanish -
// otherwise the trace does not satisfy the AIR.  a: [limb][c][n] -> h: [column][limb][n].   ce = 2, K = 1: h = (A0 + A1) / 2 as combine_kernel.
struct GoCombineConsts { u64 wi[GO_MAX_CE]; u64 scale[GO_MAX_CE]; };    // w_ce^-e;  7^(-n i) / ce
struct GoCombine {
  const u64* a; u32 ln, lce, K; int D; GoCombineConsts cc; u64* h; GoState* s;
  XFG_HD void operator()(size_t j) const {
    const size_t n = size_t(1) << ln; const u32 ce = 1u << lce; bool bad = false;
    for (int l = 0; l < D; l++) {
      u64 av[GO_MAX_CE];
      for (u32 c = 0; c < ce; c++) av[c] = a[(((size_t)l << lce) + c) * n + j];
      for (u32 i = 0; i < ce; i++) {
        u64 acc = 0;
        for (u32 c = 0; c < ce; c++) acc = gl_add(acc, gl_mul(av[c], cc.wi[(c * i) & (ce - 1)]));
        if (i < K) h[((size_t)i * D + l) * n + j] = gl_mul(acc, cc.scale[i]);
        else bad |= acc != 0;
      }
    }
    if (bad) go_flag(&s->error_flags, ERR_FLAG_DEGREE);
  }
};
// where the composition columns leave no vanishing coefficient to check (K = ce: degrees 3, 5, 9), the trace is validated directly, as winter-prover
// does in debug builds: every transition constraint on every step but the last, every assertion.  t < n - 1: step t;  t >= n - 1: assertion t - (n - 1)
struct GoValidate {
  const u64* trace; u32 ln; const GenProgram* prog; GoState* s;
  XFG_HD void operator()(size_t t) const {
    const size_t n = size_t(1) << ln;
    if (t >= n - 1) { const GenAssertion& as = prog->asr[t - (n - 1)]; const u64 step = prog->asr_step[t - (n - 1)]; if (trace[(size_t)as.column * n + step] != as.value) go_flag(&s->error_flags, ERR_FLAG_DEGREE); return; }
    u64 slot[GEN_MAX_SLOTS]; bool bad = false;
    for (u32 i = 0; i < prog->num_instr; i++) {
      const GenInstr in = prog->code[i];
      const u32 op = in.w0 & 15u, dst = in.w0 >> 8;
      u64 v[2];
      for (int o = 0; o < 2; o++) {
        const u32 kind = (in.w0 >> (4 + 2 * o)) & 3u, idx = o ? in.w1 >> 16 : in.w1 & 0xFFFFu;
        v[o] = kind == GK_SLOT ? slot[idx] : kind == GK_CONST ? prog->constants[idx] : trace[(size_t)idx * n + (kind == GK_CUR ? t : t + 1)];
        if (op == GOP_OUT) break;
      }
      if (op == GOP_OUT) { bad |= v[0] != 0; continue; }
      slot[dst] = op == GOP_MUL ? gl_mul(v[0], v[1]) : op == GOP_ADD ? gl_add(v[0], v[1]) : gl_sub(v[0], v[1]);
    }
    if (bad) go_flag(&s->error_flags, ERR_FLAG_DEGREE);
  }
};

// ------------------------------------------------------------------------------------------------------------------
// out-of-domain evaluation (A.9): polynomial p (base-field coefficients) at z and z g.  t = p * chunks + c: Horner over chunk c,
// times pt^(c L).  partial: [poly][chunk][point][GO_MAX_EXT]
// ------------------------------------------------------------------------------------------------------------------
template <int D> struct GoOodPartial {
  const u64* trace_coef; const u64* h_coef; u32 ln, width, chunks; const GoState* s; u64* partial;
  XFG_HD void operator()(size_t t) const {
    const size_t n = size_t(1) << ln, L = n / chunks, p = t / chunks, c = t % chunks;
    const u64* co = (p < width ? trace_coef + p * n : h_coef + (p - width) * n) + c * L;
    for (int w = 0; w < 2; w++) {
      const Ext<D> pt = go_ld<D>(w ? s->zg : s->z);
      Ext<D> r;
      for (size_t i = L; i-- > 0;) r = add_base(r * pt, co[i]);
      r = r * go_pow<D>(pt, (u64)(c * L));
      go_st<D>(partial + (t * 2 + w) * GO_MAX_EXT, r);
    }
  }
};
// two-level sum of the chunk partials: t = (p * 2 + point) * groups + g adds chunks g, g + groups, ... into part2; then (groups = 1, chunks = the
// first level's groups) the group sums into sums[p * 2 + point]
template <int D> struct GoOodSum {
  const u64* partial; u32 chunks, groups; u64* sums;
  XFG_HD void operator()(size_t t) const {
    const size_t pw = t / groups, g = t % groups, p = pw >> 1, w = pw & 1; Ext<D> r;
    for (u32 c = (u32)g; c < chunks; c += groups) r = r + go_ld<D>(partial + ((p * chunks + c) * 2 + w) * GO_MAX_EXT);
    go_st<D>(sums + t * GO_MAX_EXT, r);
  }
};
// second level: part2 holds `groups` consecutive group sums per (polynomial, point)
template <int D> struct GoOodSum2 {
  const u64* part2; u32 groups; u64* sums;
  XFG_HD void operator()(size_t t) const {
    Ext<D> r;
    for (u32 g = 0; g < groups; g++) r = r + go_ld<D>(part2 + (t * groups + g) * GO_MAX_EXT);
    go_st<D>(sums + t * GO_MAX_EXT, r);
  }
};
// send_ood_trace_states / send_ood_constraint_evaluations, DEEP coefficients, and the constants of the DEEP quotients (A.9)
template <int D> struct GoStepOod {
  GoState* s; const u64* sums; u32 W, K;
  XFG_HD void operator()(size_t) const {
    GoCoin c = go_coin_load(s);
    for (u32 t = 0; t < 2 * W; t++) for (int l = 0; l < GO_MAX_EXT; l++) s->ood_frame[t][l] = sums[(size_t)t * GO_MAX_EXT + l];   // sums[(2 j + w)] = T_j(z | zg): already interleaved
    const u64* fr = &s->ood_frame[0][0];
    go_reseed(c, go_hash_stream((int)(2 * W * D), [fr](int li) { return fr[(size_t)(li / D) * GO_MAX_EXT + (li % D)]; }));
    // H_i(z) = sum_l x^l P_{i,l}(z): the limb polynomials of composition column i have base-field coefficients (polynomial W + i D + l)
    for (u32 i = 0; i < K; i++) {
      Ext<D> hz;
      for (int l = D - 1; l >= 0; l--) hz = ext_mul_x<D>(hz) + go_ld<D>(sums + (size_t)(2 * (W + i * D + l)) * GO_MAX_EXT);
      go_st<D>(s->hz[i], hz);
    }
    const u64* hp = &s->hz[0][0];
    go_reseed(c, go_hash_stream((int)(K * D), [hp](int li) { return hp[(size_t)(li / D) * GO_MAX_EXT + (li % D)]; }));
    bool ok = true;
    for (u32 j = 0; j < W + K; j++) { u64 v[GO_MAX_EXT] = {0, 0, 0}; ok &= go_draw<D>(c, v); for (int l = 0; l < GO_MAX_EXT; l++) s->dcoef[j][l] = l < D ? v[l] : 0; }
    if (!ok) s->error_flags |= ERR_FLAG_COIN;
    Ext<D> c1, c2;
    for (u32 j = 0; j < W; j++) { const Ext<D> g = go_ld<D>(s->dcoef[j]); c1 = c1 + g * go_ld<D>(s->ood_frame[2 * j]); c2 = c2 + g * go_ld<D>(s->ood_frame[2 * j + 1]); }
    for (u32 i = 0; i < K; i++) c1 = c1 + go_ld<D>(s->dcoef[W + i]) * go_ld<D>(s->hz[i]);
    go_st<D>(s->deep_c1, c1); go_st<D>(s->deep_c2, c2);
    go_coin_store(s, c);
  }
};

// ------------------------------------------------------------------------------------------------------------------
// DEEP composition, pointwise (A.9; the same field elements as the reference's coefficient-domain quotients):
//   D(x) = (S_T(x) + sum_i delta_i H_i(x) - C1) / (x - z) + (S_T(x) - C2) / (x - z g),   S_T = sum_j gamma_j T_j(x)
// t = k n + m; written in natural order [limb][B m + k] = the evaluations of FRI layer 0
// ------------------------------------------------------------------------------------------------------------------
template <int D> struct GoDeep {
  const u64* lde; const u64* hlde; u32 ln, lb, W, K, pts; const GoState* s; PowTable wn; const u64* s_k; u64* deep;
  // t = k (n / pts) + q: points m = q + j n / pts, j < pts; 1 / (x - w) = adj(x - w) / N(x - w) with the 2 pts norms inverted together
  XFG_HD void operator()(size_t t) const {
    const size_t n = size_t(1) << ln, N = n << lb, per = n / pts, k = t / per, q = t % per;
    const Ext<D> z = go_ld<D>(s->z), zg = go_ld<D>(s->zg), c1 = go_ld<D>(s->deep_c1), c2 = go_ld<D>(s->deep_c2);
    Ext<D> pa[GO_PTS], qa[GO_PTS]; u64 nrm[2 * GO_PTS];
    for (u32 j = 0; j < pts; j++) {
      const size_t m = q + j * per, at = k * n + m;
      const u64 x = gl_mul(s_k[k], pow_lookup(wn, m));
      Ext<D> st;
      for (u32 c = 0; c < W; c++) st = st + mul_base(go_ld<D>(s->dcoef[c]), lde[(size_t)c * N + at]);
      Ext<D> sh = st;
      for (u32 i = 0; i < K; i++) { Ext<D> h; for (int l = 0; l < D; l++) h.set_limb(l, hlde[((size_t)i * D + l) * N + at]); sh = sh + go_ld<D>(s->dcoef[W + i]) * h; }
      const Ext<D> xe(x);
      Ext<D> az, azg;
      nrm[2 * j] = ext_norm_adj(xe - z, az); nrm[2 * j + 1] = ext_norm_adj(xe - zg, azg);
      pa[j] = (sh - c1) * az; qa[j] = (st - c2) * azg;
    }
    go_batch_inv(nrm, (int)(2 * pts));
    for (u32 j = 0; j < pts; j++) {
      const size_t m = q + j * per, i = (m << lb) | k;
      const Ext<D> r = mul_base(pa[j], nrm[2 * j]) + mul_base(qa[j], nrm[2 * j + 1]);
      for (int l = 0; l < D; l++) deep[(size_t)l * N + i] = r.limb(l);
    }
  }
};

// ------------------------------------------------------------------------------------------------------------------
// FRI (A.10): fold layer l (Nl evaluations, natural order) by F with alpha_l.  next[r] = P_r(alpha), P_r interpolating row r
// (the values at r + j R) over x_r w_F^j, x_r = 7 w_Nl^r (constant domain offset 7 at every layer):
//   P_r(alpha) = sum_k c_k (alpha / x_r)^k,   c_k = 1/F sum_j v_j w_F^(-jk)
// ------------------------------------------------------------------------------------------------------------------
struct GoFriConsts { u64 wfi[16]; u64 f_inv, inv7; };    // w_F^-j, 1/F, 1/7
template <int D> struct GoFriFold {
  const u64* src; u64 src_stride; u32 lNl, lf, layer; const GoState* s; PowTable wN_inv; u32 lN; GoFriConsts fc; u64* dst; u64 dst_stride;
  XFG_HD void operator()(size_t r) const {
    const u32 F = 1u << lf; const size_t R = (size_t(1) << lNl) >> lf;
    Ext<D> v[16];
    for (u32 j = 0; j < F; j++) for (int l = 0; l < D; l++) v[j].set_limb(l, src[(size_t)l * src_stride + r + (size_t)j * R]);
    const u64 xinv = gl_mul(fc.inv7, pow_lookup(wN_inv, (u64)r << (lN - lNl)));
    const Ext<D> beta = mul_base(go_ld<D>(s->alphas[layer]), xinv);
    Ext<D> acc;
    for (u32 kk = F; kk-- > 0;) {
      Ext<D> ck;
      for (u32 j = 0; j < F; j++) ck = ck + mul_base(v[j], fc.wfi[(j * kk) & (F - 1)]);
      acc = acc * beta + ck;
    }
    acc = mul_base(acc, fc.f_inv);
    for (int l = 0; l < D; l++) dst[(size_t)l * dst_stride + r] = acc.limb(l);
  }
};
// remainder (A.10): interpolate the last layer's Nr evaluations over the coset 7 <w_Nr> and keep the first Nr / B coefficients.
// t = j * D + l: coefficient j, limb l = 7^-j / Nr * sum_i e_i w_Nr^(-i j)
template <int D> struct GoRemainder {
  const u64* ev; u64 stride; u32 lNr, lN; PowTable wN_inv; u64 nr_inv, inv7; GoState* s;
  XFG_HD void operator()(size_t t) const {
    const size_t j = t / D, l = t % D, Nr = size_t(1) << lNr;
    u64 acc = 0;
    for (size_t i = 0; i < Nr; i++) acc = gl_add(acc, gl_mul(ev[l * stride + i], pow_lookup(wN_inv, (u64)((i * j) & (Nr - 1)) << (lN - lNr))));
    s->remainder[j][l] = gl_mul(acc, gl_mul(nr_inv, gl_pow(inv7, j)));
  }
};
// commit to the remainder polynomial (hash_elements of its coefficients); the grinding search runs on this seed
template <int D> struct GoStepRemainder {
  GoState* s; u32 rem_len;
  XFG_HD void operator()(size_t) const {
    GoCoin c = go_coin_load(s);
    s->remainder_len = rem_len;
    const u64* rp = &s->remainder[0][0];
    s->remainder_commitment = go_hash_stream((int)(rem_len * D), [rp](int li) { return rp[(size_t)(li / D) * GO_MAX_EXT + (li % D)]; });
    go_reseed(c, s->remainder_commitment);
    go_coin_store(s, c);
  }
};
// grinding (A.5): smallest nonce >= 1 with trailing_zeros(LE head of BLAKE3(seed || nonce)) >= grinding.  Thread g tests g + 1, g + 1 + TOT, ...
// and stops once its next candidate exceeds the best found so far, so the serial minimum is returned
struct GoGrind {
  GoState* s; u32 grinding; u64 total;
  XFG_HD void operator()(size_t g) const {
    const Digest seed = s->seed;
    const u64 mask = grinding >= 64 ? ~0ull : ((1ull << grinding) - 1);
    for (u64 nonce = (u64)g + 1;; nonce += total) {
      if (nonce > *(volatile unsigned long long*)&s->nonce) break;
      const Digest d = go_merge_int(seed, nonce);
      const u64 head = (u64)d.w[0] | ((u64)d.w[1] << 32);
      if ((head & mask) == 0) {
#if defined(__CUDA_ARCH__)
        atomicMin(&s->nonce, (unsigned long long)nonce);
#else
        if (nonce < s->nonce) s->nonce = nonce;
#endif
        break;
      }
    }
  }
};
// draw_integers(num_queries, N, nonce) -> sort -> dedup (A.5)
struct GoStepPositions {
  GoState* s; u32 num_queries, lN;
  XFG_HD void operator()(size_t) const {
    GoCoin c = go_coin_load(s);
    c.seed = go_merge_int(c.seed, s->nonce); c.counter = 0;
    const u64 mask = (1ull << lN) - 1;
    u32* p = s->positions; u32 cnt = 0;
    for (u32 i = 0; i < num_queries; i++) {
      c.counter += 1;
      const Digest d = go_merge_int(c.seed, c.counter);
      const u32 v = (u32)(((u64)d.w[0] | ((u64)d.w[1] << 32)) & mask);
      u32 at = cnt; while (at > 0 && p[at - 1] > v) { p[at] = p[at - 1]; at--; }     // insertion sort
      if (at > 0 && p[at - 1] == v) { for (u32 q = at; q < cnt; q++) p[q] = p[q + 1]; continue; }   // duplicate: undo the shift
      p[at] = v; cnt++;
    }
    s->num_positions = cnt;
    go_coin_store(s, c);
  }
};
// fold_positions (A.10) for layer t: p mod (N / F^(t+1)) over the query positions, first occurrence kept.  The layers are independent: a value
// dropped at one layer is a duplicate at every later one.
struct GoFoldPositions {
  GoState* s; u32 lN, lf;
  XFG_HD void operator()(size_t t) const {
    const u32 mask = (u32)((1ull << (lN - lf * ((u32)t + 1))) - 1);
    u32 cnt = 0; u32* o = s->fri_positions[t];
    for (u32 i = 0; i < s->num_positions; i++) {
      const u32 v = s->positions[i] & mask; bool dup = false;
      for (u32 q = 0; q < cnt; q++) if (o[q] == v) { dup = true; break; }
      if (!dup) o[cnt++] = v;
    }
    s->fri_num_positions[t] = cnt;
  }
};

// ------------------------------------------------------------------------------------------------------------------
// queries: opened rows and, per queried leaf, the sibling digest of every level (the host builds the BatchMerkleProofs, A.11)
// ------------------------------------------------------------------------------------------------------------------
struct GoGatherTask {
  const u64* src; const Digest* tree;
  u64 limb_stride, R, M; u32 lb, ln, coset;    // coset != 0: coset-major source (row p at (p mod B) n + p / B)
  u32 J, limbs, depth; int fri_layer;          // fri_layer < 0: the LDE query positions, else the folded positions of that layer
  u64 rows_off, paths_off;                     // offsets into the material buffer, in u64 units
  u32 max_q;
};
struct GoGather {
  GoGatherTask k; const GoState* s; u64* out;
  XFG_HD void operator()(size_t t) const {
    const u32 cnt = k.fri_layer < 0 ? s->num_positions : s->fri_num_positions[k.fri_layer];
    const u32* pos = k.fri_layer < 0 ? s->positions : s->fri_positions[k.fri_layer];
    const u32 width = k.J * k.limbs; const size_t nrows = (size_t)k.max_q * width;
    if (t < nrows) {
      const u32 q = (u32)(t / width), w = (u32)(t % width), j = w / k.limbs, l = w % k.limbs;
      if (q >= cnt) return;
      const u64 p = (u64)pos[q] + (u64)j * k.R;
      const u64 addr = k.coset ? ((p & ((1ull << k.lb) - 1)) << k.ln) + (p >> k.lb) : p;
      out[k.rows_off + t] = k.src[(size_t)l * k.limb_stride + addr];
    } else {
      const size_t e = t - nrows; const u32 q = (u32)(e / k.depth), lvl = (u32)(e % k.depth);
      if (q >= cnt) return;
      reinterpret_cast<Digest*>(out + k.paths_off)[e] = k.tree[((k.M + pos[q]) >> lvl) ^ 1];
    }
  }
};

}  // namespace xfg
