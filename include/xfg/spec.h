/* xfg/spec.h — every protocol constant and ordering decision of the burn-mint proving path, in one place.
 *
 * Shared (read-only) by the CUDA product (xfg-stark_b200/csrc) and by the CPU oracle (oracle/): both sides
 * must agree on these numbers, but no code is shared between them.
 *
 * Source of each value: SURVEY.md Appendix A (Winterfell 0.8.3 restatement; evidence level D = disassembled
 * from the shipped reference binary, P = pinned by reference source/strings, H = high-confidence recollection)
 * and the reference's own source for the AIR (src/burn_mint_air.rs, src/burn_mint_prover.rs).
 */
#ifndef XFG_SPEC_H
#define XFG_SPEC_H

#include <stdint.h>

/* ---- A.1 field: winter-math 0.8.4 fields::f64::BaseElement ---- */
#define XFG_P                 0xFFFFFFFF00000001ULL /* 2^64 - 2^32 + 1                         (D) */
#define XFG_GENERATOR         7ULL                  /* multiplicative generator = LDE offset   (D) */
#define XFG_TWO_ADICITY       32
#define XFG_TWO_ADIC_ROOT     7277203076849721926ULL /* 2^32-th root of unity                  (P) */
/* quadratic extension F_p[x]/(x^2 - x + 2): (a0,a1)(b0,b1) = (a0b0 - 2a1b1, (a0+a1)(b0+b1) - a0b0)   (D) */

/* FieldExtension discriminants (D) */
#define XFG_EXT_NONE          1
#define XFG_EXT_QUADRATIC     2
#define XFG_EXT_CUBIC         3 /* F_p[x]/(x^3 - x - 1); proved by the general-options pipeline */

/* ---- A.2 reference ProofOptions (src/burn_mint_prover.rs:28-35; real argument meaning per (D)) ---- */
#define XFG_DEF_NUM_QUERIES   42
#define XFG_DEF_BLOWUP        8
#define XFG_DEF_GRINDING      4
#define XFG_DEF_FRI_FOLDING   8
#define XFG_DEF_FRI_REM_MAX   31

/* ---- the burn-mint AIR (src/burn_mint_air.rs) ---- */
#define XFG_TRACE_WIDTH       7   /* :79-86 registers */
#define XFG_NUM_PUB_INPUTS    12  /* :54-71 ToElements order */
#define XFG_NUM_TRANSITION    7   /* :356-377 */
#define XFG_NUM_ASSERTIONS    8   /* :383-394 */
#define XFG_STD_BURN          8000000ULL      /* :208 */
#define XFG_LARGE_BURN        8000000000ULL   /* :215-216 = 8e6 * 1000 as a field element */
#define XFG_CE_BLOWUP         2   /* A.3: max(next_pow2(deg-1), 2) for declared degree 1 */
#define XFG_NUM_COMP_COLS     1   /* A.3 (the burn-mint AIR: degree 2) */
#define XFG_AIR_MAX_DEGREE    9   /* generic AIRs: highest transition-constraint degree served (ce_blowup 8, 8 composition columns) */
#define XFG_FINAL_STATE       3   /* :393 */

/* public-input slots, order of src/burn_mint_air.rs:54-71 */
enum {
  XFG_PI_BURN = 0, XFG_PI_MINT, XFG_PI_TXN_HASH, XFG_PI_RECIPIENT_HASH, XFG_PI_STATE,
  XFG_PI_TXP0, XFG_PI_TXP1, XFG_PI_TXP2, XFG_PI_TXP3, XFG_PI_NETWORK_ID, XFG_PI_TARGET_CHAIN, XFG_PI_VERSION
};

/* ---- A.5 coin ---- */
#define XFG_COIN_MAX_DRAWS    1000

/* ---- BLAKE3 (blake3 1.8.2) ---- */
#define XFG_B3_CHUNK_START    1u
#define XFG_B3_CHUNK_END      2u
#define XFG_B3_PARENT         4u
#define XFG_B3_ROOT           8u
#define XFG_DIGEST_BYTES      32

/* ---- limits that follow from the wire format (A.11/A.12) ---- */
#define XFG_MAX_QUERIES       255
#define XFG_MAX_FRI_LAYERS    16

/* ---- generic AIR front-end (xfg_air_desc): limits of this backend ---- */
#define XFG_AIR_MAX_WIDTH        128   /* one BLAKE3 chunk per row (A.6); winter-air allows 255 */
#define XFG_AIR_MAX_PUB_INPUTS   120
#define XFG_AIR_MAX_CONSTANTS    512
#define XFG_AIR_MAX_INSTR        4096
#define XFG_AIR_MAX_CONSTRAINTS  256
#define XFG_AIR_MAX_ASSERTIONS   255
#define XFG_AIR_MAX_GROUPS       16    /* distinct assertion steps = boundary-constraint divisors */
#define XFG_AIR_MAX_LIVE         64    /* simultaneously live intermediate values after register allocation */

#endif /* XFG_SPEC_H */
