#!/usr/bin/env python
"""Phase timing of the fused FRI tail kernel (debug build):  make -C xfg-stark_b200/csrc EXTRA=-DXFG_TAIL_CLOCKS OUT=../libxfgstark_dbg.so OBJDIR=../build_dbg
then  XFG_LIB=xfg-stark_b200/libxfgstark_dbg.so python tools/tail_clocks.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import xfg_stark_b200 as xs

for n_log2, ext in ((20, 2), (16, 1), (10, 2)):
    with xs.Context(device=0, max_n_log2=n_log2) as ctx:
        s = xs.synthetic_inputs(0)
        air = ctx.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
        tr = ctx.build_trace(air, n_log2)
        for _ in range(3):
            ctx.prove(tr, air, xs.ProofOptions(field_extension=ext))
        out = (C.c_longlong * 48)()
        ctx._lib.xfg_debug_tail_clocks.argtypes = [C.c_void_p, C.c_void_p]
        ctx._lib.xfg_debug_tail_clocks(ctx._h, out)
        t = [v for v in out if v]
        print(f"2^{n_log2} ext {ext}: phases (us at 1.965 GHz):", [round((b - a) / 1965.0, 1) for a, b in zip(t, t[1:])], "total", round((t[-1] - t[0]) / 1965.0, 1))
