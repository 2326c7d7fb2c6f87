"""Generic AIR front-end, CPU side: the builder's flattening, the oracle's generic prover / verifier on the
example AIRs (self-consistency, tamper rejection, unsatisfied traces), and equality of the oracle's generic path with its
hard-wired burn-mint path on the same AIR."""
import json
import os

import numpy as np
import pytest

import orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "air_proofs.json")


def examples():
    from xfg_stark_b200 import air as A
    return {"burn4": A.xfg_burn_air(0x1234567890ABCDEF, 987654321, 8_000_000, 4, 64), "fib": A.fibonacci_air(128),
            "wide12": A.wide_quadratic_air(12, 64, seed=3, extra_steps=(5, 32)), "wide40": A.wide_quadratic_air(40, 32, seed=4)}


def test_builder_flatten_and_cse():
    from xfg_stark_b200 import AirBuilder
    a = AirBuilder(3, pub_inputs=[5, orc.P + 1])
    x = a.cur(0) * a.cur(1) + 7
    y = a.cur(0) * a.cur(1) + 7          # same expression: shared instructions
    a.constraint(a.nxt(2) - x); a.constraint(a.nxt(2) - y)
    a.assert_single(2, 0, 9)
    f = a.flatten()
    assert f["desc"].tolist() == [3, 2, 1, 3, 2, 1]
    assert f["pub"].tolist() == [5, 1]
    assert f["code"].tolist() == [[2, 0, 1], [0, 7, 6], [1, 5, 8]]      # mul cur0 cur1 -> v7; add v7 const0(6) -> v8; sub nxt2(5) v8 -> v9
    assert f["outs"].tolist() == [9, 9]


def test_numpy_goldilocks_helpers():
    from xfg_stark_b200 import air as A
    rng = np.random.default_rng(0)
    a = rng.integers(0, A.P, size=500, dtype=np.uint64); b = rng.integers(0, A.P, size=500, dtype=np.uint64)
    a[:4] = [A.P - 1, A.P - 1, 0, 1 << 63]; b[:4] = [A.P - 1, 1, 5, (1 << 63) + 12345]
    m, s = A.gl_mul_np(a, b), A.gl_add_np(a, b)
    assert all(int(m[i]) == int(a[i]) * int(b[i]) % A.P and int(s[i]) == (int(a[i]) + int(b[i])) % A.P for i in range(500))


@pytest.mark.parametrize("ext", [1, 2])
@pytest.mark.parametrize("name", ["burn4", "fib", "wide12", "wide40"])
def test_oracle_generic_prove_verify(name, ext):
    air, trace = examples()[name]
    f = air.flatten(); o = (42, 8, 4, ext, 8, 31)
    proof = orc.prove_air(f, trace, o)
    assert proof[0] == air.width
    assert orc.verify_air(proof, f, o) == ""
    rng = np.random.default_rng(1)
    for pos in rng.integers(0, len(proof), size=12):
        bad = bytearray(proof); bad[pos] ^= 1 << int(rng.integers(0, 8))
        assert orc.verify_air(bytes(bad), f, o) != ""
    # a different statement (public input / assertion value) is rejected
    g = dict(f); g["asr"] = f["asr"].copy(); g["asr"][0, 2] = (int(g["asr"][0, 2]) + 1) % orc.P
    assert orc.verify_air(proof, g, o) != ""
    # golden regression pin of the oracle's generic path (tests/golden/make_golden.py)
    import hashlib
    gold = json.load(open(GOLDEN))
    assert hashlib.sha256(proof).hexdigest() == gold[f"{name}/ext{ext}"]


def test_oracle_generic_rejects_bad_traces_and_airs():
    from xfg_stark_b200 import air as A
    air, trace = A.fibonacci_air(64)
    f = air.flatten()
    t2 = trace.copy(); t2[1, 17] = (int(t2[1, 17]) + 1) % orc.P
    with pytest.raises(RuntimeError, match="UnsatisfiedTransitionConstraintError"):
        orc.prove_air(f, t2)
    g = dict(f); g["asr"] = f["asr"].copy(); g["asr"][0, 2] = 5            # wrong assertion value: boundary quotient is not a polynomial
    with pytest.raises(RuntimeError, match="UnsatisfiedTransitionConstraintError"):
        orc.prove_air(g, trace)
    b = A.AirBuilder(1); x = b.cur(0); y = x * x * x; b.constraint(b.nxt(0) - y * y * y * x); b.assert_single(0, 0, 1)
    with pytest.raises(RuntimeError, match="degree above 9"):
        orc.prove_air(b.flatten(), np.full((1, 16), 1, dtype=np.uint64))
    b = A.AirBuilder(1); b.constraint(b.nxt(0) - b.cur(0)); b.assert_single(0, 0, 2); b.assert_single(0, 0, 2)
    with pytest.raises(RuntimeError, match="duplicate assertion"):
        orc.prove_air(b.flatten(), np.full((1, 8), 2, dtype=np.uint64))
    b = A.AirBuilder(1); b.constraint(b.nxt(0) - b.cur(0))
    with pytest.raises(RuntimeError, match="at least one assertion"):
        orc.prove_air(b.flatten(), np.full((1, 8), 2, dtype=np.uint64))


@pytest.mark.parametrize("ext", [1, 2])
def test_oracle_generic_equals_hardwired_burn_mint(ext):
    """The burn-mint AIR written as a program goes through the interpreter and the generic boundary groups; the bytes must equal
    the hard-wired path's (same coefficients order, same divisors)."""
    from xfg_stark_b200 import air as A
    tr, pi, ac = orc.synthetic_case(64, 3)
    o = (42, 8, 4, ext, 8, 31)
    air = A.burn_mint_air(pi, ac[0], ac[1], ac[2], ac[3], 64)
    assert orc.prove_air(air.flatten(), tr, o) == orc.prove(tr, pi, ac, o)


# ---- the front-end's host-side compiler (validation, dead-code elimination, slot allocation), checked without a GPU ----
def test_compiler_matches_source_program_on_random_frames():
    """xfg_air_compile_check runs the COMPILED register program on one frame; it must equal the builder's own evaluation of the source
    expressions (Python integers) for programs with shared sub-expressions, dead code and many live values"""
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import air as A
    import test_gpu_air as T
    rng = np.random.default_rng(7)
    for seed in range(8):
        width = [1, 2, 3, 5, 8, 13, 21, 34][seed]
        air, trace = T.random_air(300 + seed, width, 8)
        dead = air.cur(0) * air.cur(0) + 5                      # never used by a constraint: must be eliminated
        assert dead is not None
        outs = [A.Expr(air, v) for v in air._outs]
        for _ in range(5):
            cur = rng.integers(0, orc.P, size=width, dtype=np.uint64); nxt = rng.integers(0, orc.P, size=width, dtype=np.uint64)
            r = xs.air_compile_check(air, 3, cur, nxt)
            exp = [air.evaluate(e, [int(x) for x in cur], [int(x) for x in nxt]) for e in outs]
            assert [int(x) for x in r["results"]] == exp
        assert r["num_groups"] == 2 and r["num_slots"] <= 64
        assert r["num_instr"] <= len(air._code) - 2 + len(air._outs)      # the two dead instructions are gone
    # on a satisfying trace every constraint vanishes on consecutive rows
    air, trace = A.wide_quadratic_air(9, 16, seed=5)
    for i in range(15):
        assert not xs.air_compile_check(air, 4, trace[:, i], trace[:, i + 1])["results"].any()


def test_compiler_validation_codes():
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import air as A

    def code(b, n_log2=3):
        with pytest.raises(xs.XfgError) as e:
            xs.air_compile_check(b, n_log2)
        return e.value.code
    b = A.AirBuilder(1); x = b.cur(0); b.constraint(b.nxt(0) - x * x * x); b.assert_single(0, 0, 1)
    assert xs.air_compile_check(b, 3)["num_instr"] > 0           # degree 3 compiles (two composition columns, general pipeline)
    b = A.AirBuilder(1); x = b.cur(0); y = x * x * x; b.constraint(b.nxt(0) - y * y * y * x); b.assert_single(0, 0, 1)
    assert code(b, 4) == 3                                       # degree 10: XFG_ERR_UNSUPPORTED_OPTIONS
    b = A.AirBuilder(1); x = b.cur(0); y = x * x * x; b.constraint(b.nxt(0) - y * y * y); b.assert_single(0, 0, 1)
    assert code(b, 3) == 1                                       # degree 9 on an 8-row trace: the degree must be smaller than the trace length
    b = A.AirBuilder(1); b.constraint(b.nxt(0) - b.cur(0))
    assert code(b) == 1                                          # no assertion
    b = A.AirBuilder(1); b.assert_single(0, 0, 1)
    assert code(b) == 1                                          # no constraint
    b = A.AirBuilder(1); b.constraint(b.const(5) + 1); b.assert_single(0, 0, 1)
    assert code(b) == 1                                          # degree-0 constraint
    b = A.AirBuilder(1); b.constraint(b.nxt(0) - b.cur(0)); b.assert_single(0, 0, 1); b.assert_single(0, 0, 2)
    assert code(b) == 1                                          # duplicate assertion
    b = A.AirBuilder(1); b.constraint(b.nxt(0) - b.cur(0)); b.assert_single(0, 8, 1)
    assert code(b) == 1                                          # step out of range for 2^3 rows
    b = A.AirBuilder(2); b.constraint(b.nxt(0) - b.cur(0)); b.assert_single(2, 0, 1)
    assert code(b) == 1                                          # column out of range
    b = A.AirBuilder(129); b.constraint(b.nxt(0) - b.cur(0)); b.assert_single(0, 0, 1)
    assert code(b) == 1                                          # wider than XFG_AIR_MAX_WIDTH
    b = A.AirBuilder(2); b.constraint(b.nxt(0) - b.cur(0))
    for s in range(17):
        b.assert_single(0, s, 1)
    assert code(b, 5) == 3                                       # 17 distinct assertion steps
    b = A.AirBuilder(2); vals = [b.cur(0) + (i + 1) for i in range(80)]
    acc = vals[0] * b.cur(1)
    for v in vals[1:]:
        acc = acc + v * b.cur(1)
    b.constraint(b.nxt(1) - acc); b.assert_single(0, 0, 1)
    assert code(b) == 3                                          # 80 values alive at once > XFG_AIR_MAX_LIVE
    ok = A.AirBuilder(2); ok.constraint(ok.nxt(0) - ok.cur(0) * ok.cur(1)); ok.constraint(ok.cur(1) - 3); ok.assert_single(0, 0, 1); ok.assert_single(1, 7, 3)
    r = xs.air_compile_check(ok, 3)
    assert r["num_groups"] == 2 and r["num_instr"] == 5 and r["num_slots"] >= 1       # mul, sub, OUT, sub, OUT


def test_compiler_survives_hostile_descriptions():
    """random, mostly malformed xfg_air_desc arrays (indices out of range, absurd counts, non-canonical values) must come back with an error
    code or a valid compilation - never a crash; whatever compiles must evaluate like the straight-line source"""
    import ctypes as C
    import xfg_stark_b200 as xs
    from xfg_stark_b200._binding import _AirDesc, _Assertion, load_library
    L = load_library()
    rng = np.random.default_rng(11)
    ok = bad = 0
    for _ in range(400):
        w = int(rng.integers(0, 140)); nc = int(rng.integers(0, 6)); ni = int(rng.integers(0, 40)); no = int(rng.integers(0, 6)); na = int(rng.integers(0, 6))
        hi = 2 * max(w, 1) + nc + ni + 3
        consts = rng.integers(0, 1 << 64, size=max(nc, 1), dtype=np.uint64, endpoint=False) if rng.integers(0, 4) == 0 else rng.integers(0, orc.P, size=max(nc, 1), dtype=np.uint64)
        code = np.stack([rng.integers(0, 4, size=max(ni, 1)), rng.integers(0, hi, size=max(ni, 1)), rng.integers(0, hi, size=max(ni, 1))], axis=1).astype(np.uint32)
        outs = rng.integers(0, hi, size=max(no, 1)).astype(np.uint32)
        asr = (_Assertion * max(na, 1))(*[_Assertion(int(rng.integers(0, max(w, 1) + 2)), int(rng.integers(0, 20)), int(rng.integers(0, orc.P, dtype=np.uint64))) for _ in range(max(na, 1))])
        if _ % 2 and 1 <= w <= 128 and ni and no and na:      # every other case: structurally valid (operands refer to earlier values, mostly linear operations)
            first = 2 * w + nc
            code = np.array([[int(rng.choice([0, 0, 1, 1, 2])), int(rng.integers(0, first + i)), int(rng.integers(0, first + i))] for i in range(ni)], dtype=np.uint32)
            outs = rng.integers(first, first + ni, size=no).astype(np.uint32)
            asr = (_Assertion * na)(*[_Assertion(int(rng.integers(0, w)), k, int(rng.integers(0, orc.P, dtype=np.uint64))) for k in range(na)])
        pub = rng.integers(0, orc.P, size=3, dtype=np.uint64)
        d = _AirDesc(w, 3, nc, ni, no, na, pub.ctypes.data, consts.ctypes.data, code.ctypes.data, outs.ctypes.data, C.addressof(asr))
        n1, n2, n3 = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        cur = rng.integers(0, orc.P, size=max(w, 1), dtype=np.uint64); nxt = rng.integers(0, orc.P, size=max(w, 1), dtype=np.uint64)
        res = np.zeros(max(no, 1), dtype=np.uint64)
        rc = L.xfg_air_compile_check(C.byref(d), 4, C.byref(n1), C.byref(n2), C.byref(n3), cur.ctypes.data, nxt.ctypes.data, res.ctypes.data)
        assert rc in (0, 1, 3)
        if rc:
            bad += 1
            continue
        ok += 1
        vals = [int(x) for x in cur] + [int(x) for x in nxt] + [int(x) for x in consts[:nc]]
        for op, a, b in code[:ni]:
            x, y = vals[int(a)], vals[int(b)]
            vals.append((x + y) % orc.P if op == 0 else (x - y) % orc.P if op == 1 else x * y % orc.P)
        assert [int(v) for v in res[:no]] == [vals[int(o)] for o in outs[:no]]
        assert n1.value <= ni + no and n2.value <= 64 and 1 <= n3.value <= 16
    assert bad > 150 and ok > 30, (ok, bad)
