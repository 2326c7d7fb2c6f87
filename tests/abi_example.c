/* Plain-C consumer of include/xfg_stark.h (what a cgo / FFI binding sees): compiled and run by tests/test_abi.py on a CPU-only box.
 * It exercises the host-only entry points and the no-CPU-fallback rule; with a GPU present it also runs one small generic-AIR proof. */
#include <stdio.h>
#include <string.h>
#include "xfg_stark.h"

int main(void) {
  /* the XfgBurnAir sketch (src/winterfell_air.rs:87-127): result[i] = current[i] - expected_i, Assertion::single(i, 0, expected_i) */
  uint64_t consts[4] = {11, 22, 8000000, 4};
  xfg_air_instr code[4]; uint32_t outs[4]; xfg_assertion asr[4];
  for (uint32_t i = 0; i < 4; i++) {
    code[i].op = XFG_OP_SUB; code[i].a = i; code[i].b = 2 * 4 + i;      /* current[i] - constants[i] */
    outs[i] = 2 * 4 + 4 + i;
    asr[i].column = i; asr[i].step = 0; asr[i].value = consts[i];
  }
  xfg_air_desc air; memset(&air, 0, sizeof air);
  air.width = 4; air.num_constants = 4; air.num_instr = 4; air.num_constraints = 4; air.num_assertions = 4;
  air.constants = consts; air.code = code; air.constraint_values = outs; air.assertions = asr;
  uint32_t ni = 0, ns = 0, ng = 0; uint64_t cur[4] = {11, 22, 8000000, 5}, res[4];
  int rc = xfg_air_compile_check(&air, 6, &ni, &ns, &ng, cur, cur, res);
  if (rc != XFG_OK || ni != 8 || ng != 1 || res[0] != 0 || res[3] != 1) { printf("compile check failed: rc %d ni %u ng %u\n", rc, ni, ng); return 1; }
  air.num_assertions = 0;
  if (xfg_air_compile_check(&air, 6, 0, 0, 0, 0, 0, 0) != XFG_ERR_BAD_ARGS) { printf("missing assertions not rejected\n"); return 1; }
  air.num_assertions = 4;
  if (strcmp(xfg_strerror(XFG_ERR_UNSATISFIED_CONSTRAINT), "UnsatisfiedTransitionConstraintError") != 0) return 1;

  xfg_ctx* ctx = 0;
  rc = xfg_create(0, 6, 1, &ctx);
  if (rc == XFG_ERR_CUDA) { printf("no CUDA device: xfg_create refused (no CPU fallback)\nABI_EXAMPLE_OK\n"); return 0; }
  if (rc != XFG_OK) { printf("xfg_create: %s\n", xfg_strerror(rc)); return 1; }
  static uint64_t trace[4 * 64]; static uint8_t proof[1 << 16]; size_t len = 0;
  for (int c = 0; c < 4; c++) for (int i = 0; i < 64; i++) trace[c * 64 + i] = consts[c];
  xfg_options opt = {XFG_DEF_NUM_QUERIES, XFG_DEF_BLOWUP, XFG_DEF_GRINDING, XFG_EXT_NONE, XFG_DEF_FRI_FOLDING, XFG_DEF_FRI_REM_MAX};
  rc = xfg_prove_air(ctx, &air, trace, 6, &opt, proof, sizeof proof, &len, 0);
  printf("xfg_prove_air: %s, %zu proof bytes\n", xfg_strerror(rc), len);
  xfg_destroy(ctx);
  if (rc != XFG_OK || len < 1000 || proof[0] != 4) return 1;
  printf("ABI_EXAMPLE_OK\n");
  return 0;
}
