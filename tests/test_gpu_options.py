"""GPU parity of the general-options pipeline (`XfgBurnMintProver::with_options`, src/burn_mint_prover.rs:44-49): blowup factors 2..128, FRI
folding factors 2/4/8/16, remainder degrees 0..255 and all three field extensions, through the C ABI.

Checked against (1) the proofs the REFERENCE's own Winterfell 0.8.3 prover emitted for such options (tests/golden/reference_proofs_options.json,
oracle/a64emu/make_reference_option_vectors.py) and (2) the oracle on seeded sweeps, larger traces and wide generic AIRs."""
import random

import numpy as np
import pytest

import orc
import refvec

pytestmark = pytest.mark.gpu


def _opts(xs, o):
    return xs.ProofOptions(num_queries=o[0], blowup_factor=o[1], grinding_factor=o[2], field_extension=o[3], fri_folding_factor=o[4], fri_remainder_max_degree=o[5])


@pytest.fixture(scope="module")
def ctx():
    import xfg_stark_b200 as xs
    with xs.Context(device=0, max_n_log2=17, num_slots=2, max_width=16) as c:
        yield c


@pytest.mark.parametrize("name", refvec.option_case_ids())
def test_gpu_proof_equals_reference_proof_for_other_options(ctx, name):
    import xfg_stark_b200 as xs
    c = next(x for x in refvec.option_cases() if x["name"] == name)
    ref = refvec.proof_bytes(c)
    pi, ac, o, n = refvec.statement(c)
    opts = _opts(xs, o)
    t = refvec.trace(c, pi, ac)
    # generic AIR front-end with the source's literal assertion step (src/burn_mint_air.rs:393): any trace length
    assert ctx.prove_air(refvec.air_program(c, pi, ac), t, opts) == ref
    if refvec.is_normalised(c):
        # 64 rows: the burn-mint entry points of the reference's path
        args = (8_000_000, 8_000_000, bytes.fromhex(c["tx_prefix_hash"]), bytes.fromhex(c["recipient"]), bytes.fromhex(c["secret"]), c["network_id"], c["target_chain_id"], c["version"])
        air = xs.pack_inputs(*args)
        assert ctx.prove(t, air, opts) == ref
        assert ctx.prove_from_inputs(*args, n_log2=6, options=opts) == ref
        R = (1 << 64) % orc.P
        mont = [np.array([(int(v) * R) % orc.P for v in t[k]], dtype=np.uint64) for k in range(7)]     # TraceTable memory: Montgomery form
        assert ctx.prove_cols(mont, air, opts, form=1) == ref


def test_seeded_option_sweep_equals_oracle(ctx):
    import xfg_stark_b200 as xs
    rng = random.Random(20261020)
    done = 0
    while done < 30:
        n_log2 = rng.choice([3, 4, 5, 6, 7, 8, 9, 10, 11, 12])
        blowup = rng.choice([2, 4, 8, 16, 32, 64, 128]); folding = rng.choice([2, 4, 8, 16]); rem = rng.choice([0, 1, 3, 7, 15, 31, 63, 127, 255])
        if n_log2 + blowup.bit_length() - 1 > 17:
            continue
        o = (rng.randrange(1, min(255, (blowup << n_log2) - 1) + 1), blowup, rng.randrange(0, 13), rng.choice([1, 2, 3]), folding, rem)
        index = rng.randrange(1 << 20)
        tr, pi, ac = orc.synthetic_case(1 << n_log2, index)
        s = orc.synthetic_inputs(index)
        air = xs.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
        if refvec.fri_shape_refused(n_log2, o):
            with pytest.raises(xs.XfgError) as e:
                ctx.prove(tr, air, _opts(xs, o))
            assert e.value.code == 2, (n_log2, o)                # refused where the reference itself panics
            continue
        assert ctx.prove(tr, air, _opts(xs, o)) == orc.prove(tr, pi, ac, o), (n_log2, o)
        done += 1


@pytest.mark.parametrize("n_log2,o", [(16, (42, 16, 4, 3, 4, 15)), (17, (42, 4, 8, 2, 16, 7)), (15, (60, 2, 4, 1, 2, 0)), (13, (42, 128, 4, 3, 8, 255))])
def test_longer_traces_equal_oracle(n_log2, o):
    """2^16 / 2^17 rows: four-step NTTs through the two-level twiddle lookups (the general plans keep no direct tables), 2^20-point LDE domains"""
    import xfg_stark_b200 as xs
    tr, pi, ac = orc.synthetic_case(1 << n_log2, 11)
    s = orc.synthetic_inputs(11)
    air = xs.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
    orc.set_threads(orc.max_threads())
    expect = orc.prove(tr, pi, ac, o)
    orc.set_threads(1)
    with xs.Context(device=0, max_n_log2=min(24, n_log2 + 5), num_slots=1) as c:
        proof, times = c.prove(tr, air, _opts(xs, o), want_times=True)
        assert proof == expect
        assert times["kernel_launches"] > 20
        import torch
        dev = torch.from_numpy(np.ascontiguousarray(tr).view(np.int64)).cuda()
        assert c.prove_device(dev.data_ptr(), n_log2, air, _opts(xs, o)) == expect
    assert orc.verify(proof, pi, ac, o) == ""


def test_wide_generic_airs_with_other_options_equal_oracle():
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import air as A
    with xs.Context(device=0, max_n_log2=14, num_slots=1, max_width=128) as c:
        for width, n, o in [(33, 512, (42, 4, 4, 3, 4, 7)), (128, 256, (30, 16, 2, 3, 2, 3)), (9, 1024, (42, 32, 4, 2, 16, 31)), (2, 64, (17, 2, 0, 1, 2, 1))]:
            air, trace = A.wide_quadratic_air(width, n, seed=width)
            assert c.prove_air(air, trace, _opts(xs, o)) == orc.prove_air(air.flatten(), trace, o), (width, n, o)


def test_batches_and_errors_with_other_options(ctx):
    import xfg_stark_b200 as xs
    o = (42, 4, 4, 3, 4, 3)
    cases = [orc.synthetic_case(256, i) for i in range(3)]
    airs = []
    for i in range(3):
        s = orc.synthetic_inputs(i)
        airs.append(xs.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"]))
    proofs, _ = ctx.prove_batch([t for t, _, _ in cases], airs, _opts(xs, o))
    for i, (t, pi, ac) in enumerate(cases):
        assert proofs[i] == orc.prove(t, pi, ac, o)
    bad = cases[1][0].copy(); bad[4, 17] = 2                      # unsatisfied state-transition constraint
    with pytest.raises(xs.XfgError) as e:
        ctx.prove(bad, airs[1], _opts(xs, o))
    assert e.value.code == 5
    bad = cases[1][0].copy(); bad[2, 5] = orc.P + 1               # non-canonical element
    with pytest.raises(xs.XfgError) as e:
        ctx.prove(bad, airs[1], _opts(xs, o))
    assert e.value.code == 1
    with xs.Context(device=0, max_n_log2=8, num_slots=1) as small:    # slot workspaces are sized for blowup 8: a blowup-128 proof at the context's maximum length gets its own
        o128 = (42, 128, 4, 3, 8, 31)
        assert small.prove(cases[0][0], airs[0], _opts(xs, o128)) == orc.prove(cases[0][0], cases[0][1], cases[0][2], o128)
        assert small.prove(cases[0][0], airs[0], _opts(xs, o)) == proofs[0]
        assert small.prove(cases[0][0], airs[0]) == orc.prove(*cases[0])
    # the tuned pipeline is untouched by a general proof on the same context
    t, pi, ac = cases[2]
    assert ctx.prove(t, airs[2]) == orc.prove(t, pi, ac)
    # batch verification with these options: the general verifier (one thread per proof)
    assert ctx.verify_batch(proofs, airs, _opts(xs, o)) == ["", "", ""]
    assert ctx.verify_batch([proofs[0]], [airs[1]], _opts(xs, o)) != [""]


@pytest.mark.parametrize("name", refvec.degree_case_ids())
def test_gpu_proof_equals_reference_proof_with_declared_degrees(ctx, name):
    """AIRs declaring transition degrees 3 .. 9 (d - 1 composition columns): bytes of the reference binary"""
    import xfg_stark_b200 as xs
    c = next(x for x in refvec.degree_cases() if x["name"] == name)
    pi, ac, o, n = refvec.statement(c)
    assert ctx.prove_air(refvec.air_program(c, pi, ac), refvec.trace(c, pi, ac), _opts(xs, o)) == refvec.proof_bytes(c)


def test_real_higher_degree_airs_equal_oracle(ctx):
    """constraints of actual degree 3 .. 9: every composition column non-zero; default options route here as well (several columns)"""
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import air as A
    for d, o, w, n in [(3, (42, 8, 4, 2, 8, 31), 3, 256), (4, (42, 8, 4, 1, 8, 31), 2, 4096), (5, (30, 4, 2, 3, 4, 7), 3, 256), (6, (42, 8, 4, 2, 8, 31), 1, 2048),
                       (9, (42, 16, 4, 2, 8, 31), 2, 128), (3, (20, 2, 0, 1, 2, 0), 4, 64), (7, (33, 8, 3, 3, 2, 15), 2, 512), (3, (42, 8, 4, 2, 8, 31), 5, 1 << 14)]:
        air, t = A.power_map_air(w, n, d, seed=d)
        orc.set_threads(orc.max_threads())
        expect = orc.prove_air(air.flatten(), t, o)
        orc.set_threads(1)
        assert ctx.prove_air(air, t, _opts(xs, o)) == expect, (d, o)
        bad = t.copy(); bad[0, n // 3] ^= 1
        with pytest.raises(xs.XfgError) as e:
            ctx.prove_air(air, bad, _opts(xs, o))
        assert e.value.code == 5, (d, o)
    air, t = A.power_map_air(2, 64, 5, seed=1)
    with pytest.raises(xs.XfgError) as e:
        ctx.prove_air(air, t, _opts(xs, (20, 2, 0, 1, 2, 1)))       # degree 5 needs a constraint-evaluation blowup of 4
    assert e.value.code == 2
    proofs, _ = ctx.prove_air_batch([A.power_map_air(2, 128, 3, seed=9)[0], A.fibonacci_air(128)[0]], [A.power_map_air(2, 128, 3, seed=9)[1], A.fibonacci_air(128)[1]])
    assert proofs[0] == orc.prove_air(A.power_map_air(2, 128, 3, seed=9)[0].flatten(), A.power_map_air(2, 128, 3, seed=9)[1], (42, 8, 4, 1, 8, 31))
    assert proofs[1] == orc.prove_air(A.fibonacci_air(128)[0].flatten(), A.fibonacci_air(128)[1], (42, 8, 4, 1, 8, 31))


def test_general_verifier_gives_the_oracle_verdicts(ctx):
    """xfg_verify_air_batch / xfg_verify_burn_mint_batch outside the tuned option set: every tampered proof gets the oracle verifier's verdict"""
    import xfg_stark_b200 as xs
    import test_options_pins as T
    from xfg_stark_b200 import air as A
    builders = {"burn-mint": None}
    t, pi, ac = orc.synthetic_case(64, 1)
    s = orc.synthetic_inputs(1)
    bm_consts = xs.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
    bm = A.burn_mint_air(pi, ac[0], ac[1], ac[2], ac[3], 64)
    a3, t3 = A.power_map_air(2, 256, 3, seed=3); a5, t5 = A.power_map_air(2, 128, 5, seed=5); w9, tw = A.wide_quadratic_air(9, 512, seed=4)
    cases = [("burn-mint cubic folding 4", bm, t, (30, 16, 3, 3, 4, 7)), ("burn-mint blowup 2 folding 2", bm, t, (20, 2, 0, 2, 2, 0)), ("degree 3", a3, t3, (42, 8, 4, 2, 8, 31)),
             ("degree 5 cubic folding 16", a5, t5, (25, 4, 2, 3, 16, 7)), ("nine columns", w9, tw, (42, 8, 4, 1, 8, 31))]
    for name, air, tr, o in cases:
        f = air.flatten()
        proof = ctx.prove_air(air, tr, _opts(xs, o))
        assert proof == orc.prove_air(f, tr, o)
        bad = T.tampered(proof, 150) + [proof]
        ours = ctx.verify_air_batch(bad, [air] * len(bad), _opts(xs, o))
        theirs = [orc.verify_air(b, f, o) for b in bad]
        assert ours[-1] == "" and theirs[-1] == "", name
        mism = [(i, a, b) for i, (a, b) in enumerate(zip(ours, theirs)) if not T.same_verdict(a, b)]
        assert not mism, (name, mism[:5])
        assert len({a for a, b in zip(ours, theirs) if a in T.CRYPTO and a == b}) >= 5
        if name.startswith("burn-mint"):      # the burn-mint entry point routes here for these options
            res = ctx.verify_batch(bad, [bm_consts] * len(bad), _opts(xs, o))
            assert res == ours, name
    # the reference binary's own proofs for other options and declared degrees
    for c in refvec.option_cases()[:8] + refvec.degree_cases()[:6]:
        rpi, rac, ro, rn = refvec.statement(c)
        assert ctx.verify_air_batch([refvec.proof_bytes(c)], [refvec.air_program(c, rpi, rac)], _opts(xs, ro)) == [""], c["name"]
    # a batch of 256 proofs at once (one thread each)
    air, tr, o = bm, t, (30, 16, 3, 3, 4, 7)
    proof = ctx.prove_air(air, tr, _opts(xs, o))
    res, times = ctx.verify_air_batch([proof] * 256, [air] * 256, _opts(xs, o), want_times=True)
    assert res == [""] * 256 and times["kernel_ms"] > 0


def test_full_size_round_trips_and_verifier_cross_check():
    """BASELINE's full trace length (2^20 rows) through the general-options pipeline: prove -> verify round trips (size-independent property; the oracle
    PROVER needs seconds per case at this size, its verifier milliseconds), and the two GPU verifiers against each other on a tuned-pipeline proof"""
    import torch
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import air as A
    n_log2 = 20
    s = orc.synthetic_inputs(3)
    consts = xs.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
    tr, pi, ac = orc.synthetic_case(1 << n_log2, 3)
    bm = A.burn_mint_air(pi, ac[0], ac[1], ac[2], ac[3], 1 << n_log2)
    with xs.Context(device=0, max_n_log2=n_log2, num_slots=1) as c:
        dev = torch.from_numpy(np.ascontiguousarray(tr).view(np.int64)).cuda()
        for o in [(42, 8, 4, 2, 4, 31), (42, 8, 4, 3, 8, 31), (42, 4, 4, 1, 16, 7)]:
            proof = c.prove_device(dev.data_ptr(), n_log2, consts, _opts(xs, o))
            assert orc.verify(proof, pi, ac, o) == "", o                                   # the oracle verifier accepts
            assert c.verify_batch([proof], [consts], _opts(xs, o)) == [""], o              # ... and so does the general GPU verifier
            bad = bytearray(proof); bad[len(bad) // 3] ^= 4
            assert c.verify_batch([bytes(bad)], [consts], _opts(xs, o)) != [""]
        tuned = c.prove_device(dev.data_ptr(), n_log2, consts, _opts(xs, (42, 8, 4, 2, 8, 31)))
        assert c.verify_batch([tuned], [consts], _opts(xs, (42, 8, 4, 2, 8, 31))) == [""]  # block-per-proof verifier (verify.cu)
        assert c.verify_air_batch([tuned], [bm], _opts(xs, (42, 8, 4, 2, 8, 31))) == [""]  # thread-per-proof verifier on the same bytes
