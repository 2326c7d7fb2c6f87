"""Memory-safety check without compute-sanitizer (closed on this GPU pool): every region of a proof's workspace is followed by a guard zone;
the slab is painted, proofs run through every entry point, and no guard word - nor the slack behind the last region - may have changed.
An out-of-bounds store of any kernel into a neighbouring region's border shows up here; the proof bytes are compared with the oracle as usual."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_log2,ext,rem", [(3, 1, 31), (4, 2, 7), (6, 2, 31), (9, 1, 15), (11, 2, 31), (13, 1, 31), (16, 2, 31), (17, 1, 63), (18, 2, 31)])
def test_guard_zones_survive_burn_mint_proofs(n_log2, ext, rem):
    import xfg_stark_b200 as xs
    opts = xs.ProofOptions(field_extension=ext, fri_remainder_max_degree=rem)
    s = orc.synthetic_inputs(n_log2)
    with xs.Context(device=0, max_n_log2=n_log2) as ctx:
        air = ctx.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
        trace = ctx.build_trace(air, n_log2)
        ctx.guard_fill()
        p1 = ctx.prove(trace, air, opts)                                  # captured launch sequence
        p2, _ = ctx.prove(trace, air, opts, want_times=True)              # launch by launch
        p3 = ctx.prove(trace, air, opts)                                  # graph replay
        p4 = ctx.prove_from_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"],
                                   n_log2=n_log2, options=opts)
        assert p1 == p2 == p3 == p4
        assert ctx.guard_check(n_log2, opts) == (0, -1)
    tr, pi, ac = orc.synthetic_case(1 << n_log2, n_log2)
    assert p1 == orc.prove(tr, pi, ac, opts.as_tuple())


@pytest.mark.parametrize("width,n_log2,ext", [(1, 5, 1), (2, 8, 2), (9, 10, 2), (33, 12, 1), (128, 9, 2)])
def test_guard_zones_survive_generic_air_proofs(width, n_log2, ext):
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import air as A
    opts = xs.ProofOptions(field_extension=ext)
    air, trace = A.wide_quadratic_air(width, 1 << n_log2, seed=width)
    with xs.Context(device=0, max_n_log2=n_log2, max_width=width) as ctx:
        ctx.guard_fill()
        p1 = ctx.prove_air(air, trace, opts); p2 = ctx.prove_air(air, trace, opts)
        assert p1 == p2 == orc.prove_air(air.flatten(), trace, opts.as_tuple())
        assert ctx.guard_check(n_log2, opts, width=width) == (0, -1)


def test_guard_check_detects_a_stray_store():
    """the checker itself: a deliberate out-of-bounds write (one word behind the trace region, through the device-pointer entry) is reported"""
    import torch
    import xfg_stark_b200 as xs
    n_log2 = 8
    with xs.Context(device=0, max_n_log2=n_log2) as ctx:
        ctx.guard_fill()
        assert ctx.guard_check(n_log2) == (0, -1)
        import ctypes as C
        ctx._lib.xfg_debug_poke_guard.argtypes = [C.c_void_p, C.c_uint32]
        assert ctx._lib.xfg_debug_poke_guard(ctx._h, n_log2) == 0
        bad, region = ctx.guard_check(n_log2)
        assert bad == 1 and region == 0
