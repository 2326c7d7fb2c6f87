"""`ProofOptions` generality pinned against the REFERENCE ITSELF (CPU tests).

`XfgBurnMintProver::with_options` (src/burn_mint_prover.rs:44-49) accepts every `ProofOptions::new` value.  The reference's own Winterfell
0.8.3 prover (executed from its shipped binary by oracle/a64emu) emitted tests/golden/reference_proofs_options.json for blowup factors
2..128, FRI folding factors 2/4/8/16, remainder degrees 0..255 and all three field extensions.  Checked here, on any box:

  1. the oracle's bytes equal the reference's on every case (this is what pins oracle/field.hpp F3, the cubic extension);
  2. the PRODUCT's general-options pipeline - its launch sequence and per-thread kernel bodies (xfg-stark_b200/csrc/general_*.cuh), executed
     on the host by tests/host_emul - emits the same bytes; the GPU runs the same bodies as CUDA kernels (tests/test_gpu_options.py);
  3. option sets the reference itself refuses with a panic are refused with an error code.
"""
import hashlib
import random

import numpy as np
import pytest

import goemul
import orc
import refvec


@pytest.mark.parametrize("name", refvec.option_case_ids())
def test_oracle_and_emulated_pipeline_equal_reference_proof(name):
    c = next(x for x in refvec.option_cases() if x["name"] == name)
    ref = refvec.proof_bytes(c)
    assert hashlib.sha256(ref).hexdigest() == c["proof_sha256"] and len(ref) == c["proof_len"]
    pi, ac, o, n = refvec.statement(c)
    t = refvec.trace(c, pi, ac)
    air = refvec.air_program(c, pi, ac).flatten()
    assert orc.prove_air(air, t, o) == ref
    assert orc.verify_air(ref, air, o) == ""
    assert goemul.prove_air(air, t, o) == ref
    if refvec.is_normalised(c):                                  # 64 rows: the normalised AIR (last step n - 1) IS the source's AIR
        assert orc.prove(t, pi, ac, o) == ref
        assert orc.verify(ref, pi, ac, o) == ""
        assert goemul.prove_burn_mint(t, pi, ac, o) == ref       # the C++ AIR description the product uses for the burn-mint entry points
        R = (1 << 64) % orc.P
        mont = np.array([[(int(v) * R) % orc.P for v in row] for row in t], dtype=np.uint64)
        assert goemul.prove_burn_mint(mont, pi, ac, o, montgomery=True) == ref


@pytest.mark.parametrize("name", refvec.degree_case_ids())
def test_oracle_and_emulated_pipeline_equal_reference_proof_with_declared_degrees(name):
    """AIRs declaring transition degrees 3 .. 9: d - 1 composition columns, constraint-evaluation blowup next_pow2(d - 1).  The program writes constraint 1
    as an expression of that degree with the same values (c0^d - c0^d), which is what the reference computes under the declared degree."""
    c = next(x for x in refvec.degree_cases() if x["name"] == name)
    ref = refvec.proof_bytes(c)
    assert hashlib.sha256(ref).hexdigest() == c["proof_sha256"]
    pi, ac, o, n = refvec.statement(c)
    t = refvec.trace(c, pi, ac)
    air = refvec.air_program(c, pi, ac).flatten()
    assert orc.prove_air(air, t, o) == ref
    assert orc.verify_air(ref, air, o) == ""
    assert goemul.prove_air(air, t, o) == ref


def test_real_higher_degree_airs_emulated_pipeline_equals_oracle():
    """constraints of ACTUAL degree 3 .. 9 (every composition column non-zero): the oracle's verifier accepts the oracle's proof (the OOD check recombines
    the columns as sum_i z^(i n) H_i(z)), the emulated product pipeline emits the same bytes, and a violated trace is refused - also where the columns
    leave no vanishing coefficient to check (degrees 3, 5, 9: validated on the trace itself)"""
    from xfg_stark_b200 import air as A
    for d, o, w, n in [(3, (42, 8, 4, 2, 8, 31), 3, 256), (4, (42, 8, 4, 1, 8, 31), 2, 1024), (5, (30, 4, 2, 3, 4, 7), 3, 256), (6, (42, 8, 4, 2, 8, 31), 1, 2048),
                       (9, (42, 16, 4, 2, 8, 31), 2, 128), (3, (20, 2, 0, 1, 2, 0), 4, 64), (7, (33, 8, 3, 3, 2, 15), 2, 512), (8, (42, 128, 0, 1, 16, 255), 1, 64)]:
        air, t = A.power_map_air(w, n, d, seed=d)
        f = air.flatten()
        expect = orc.prove_air(f, t, o)
        assert orc.verify_air(expect, f, o) == "", (d, o)
        assert goemul.prove_air(f, t, o) == expect, (d, o)
        bad = t.copy(); bad[0, n // 3] ^= 1
        with pytest.raises(goemul.EmulError) as e:
            goemul.prove_air(f, bad, o)
        assert e.value.code == 5, (d, o)
    air, t = A.power_map_air(2, 64, 5, seed=1)                    # degree 5 needs a constraint-evaluation blowup of 4: refused with blowup 2
    with pytest.raises(goemul.EmulError) as e:
        goemul.prove_air(air.flatten(), t, (20, 2, 0, 1, 2, 1))
    assert e.value.code == 2


CRYPTO = {"UnacceptableProofOptions", "InconsistentOodConstraintEvaluations", "QuerySeedProofOfWorkVerificationFailed", "NumberOfQueriesMismatch",
          "TraceQueryDoesNotMatchCommitment", "ConstraintQueryDoesNotMatchCommitment", "LayerCommitmentMismatch", "InvalidLayerFolding",
          "RemainderCommitmentMismatch", "RemainderDegreeMismatch", "InvalidRemainderFolding", "DegreeTruncation"}


def same_verdict(ours, oracle):
    """as tests/test_gpu_verify.py: parse-level rejections are one code on the product side"""
    if oracle == "":
        return ours == ""
    if ours == "":
        return False
    return ours == "ProofDeserializationError" or (oracle in CRYPTO and ours == oracle)


def verifier_cases():
    from xfg_stark_b200 import air as A
    t, pi, ac = orc.synthetic_case(64, 1)
    bm = A.burn_mint_air(pi, ac[0], ac[1], ac[2], ac[3], 64).flatten()
    a3, t3 = A.power_map_air(2, 256, 3, seed=3); a5, t5 = A.power_map_air(2, 128, 5, seed=5); w9, tw = A.wide_quadratic_air(9, 512, seed=4)
    return [("burn-mint default", bm, t, (42, 8, 4, 1, 8, 31)), ("burn-mint cubic folding 4", bm, t, (30, 16, 3, 3, 4, 7)), ("burn-mint blowup 2 folding 2", bm, t, (20, 2, 0, 2, 2, 0)),
            ("degree 3", a3.flatten(), t3, (42, 8, 4, 2, 8, 31)), ("degree 5 cubic folding 16", a5.flatten(), t5, (25, 4, 2, 3, 16, 7)), ("nine columns", w9.flatten(), tw, (42, 8, 4, 1, 8, 31))]


def tampered(proof, count=120):
    rng = np.random.default_rng(len(proof))
    offs = sorted(set(range(0, 64)) | set(range(len(proof) - 48, len(proof))) | {int(x) for x in rng.integers(0, len(proof), size=count)})
    bad = []
    for off in offs:
        b = bytearray(proof); b[off] ^= 1 << int(rng.integers(0, 8)); bad.append(bytes(b))
    return bad + [proof[:-1], proof + b"\0", proof[:len(proof) // 2], b""]


def test_emulated_general_verifier_gives_the_oracle_verdicts():
    """the product's one-thread-per-proof verifier body (general_verify.cuh) run on the host: accepts what the oracle verifier accepts and names the same
    failing check on ~240 tampered proofs per case (options incl. cubic / folding 2..16, multi-column AIRs)"""
    for name, f, tr, o in verifier_cases():
        proof = orc.prove_air(f, tr, o)
        assert orc.verify_air(proof, f, o) == "" and goemul.verify_air(f, proof, o) == "", name
        fired = set()
        for b in tampered(proof):
            ours, theirs = goemul.verify_air(f, b, o), orc.verify_air(b, f, o)
            assert ours != "" and same_verdict(ours, theirs), (name, ours, theirs)
            if ours in CRYPTO and ours == theirs:
                fired.add(ours)
        assert len(fired) >= 5, (name, fired)
        other = list(o); other[0] += 1
        assert goemul.verify_air(f, proof, tuple(other)) == "UnacceptableProofOptions" == orc.verify_air(proof, f, tuple(other))


def test_emulated_general_verifier_accepts_the_reference_binaries_proofs():
    for c in refvec.option_cases() + refvec.degree_cases() + refvec.cases()[:4]:
        pi, ac, o, n = refvec.statement(c)
        air = refvec.air_program(c, pi, ac).flatten()
        assert goemul.verify_air(air, refvec.proof_bytes(c), o) == "", c["name"]
    # a shape both verifiers reject although the prover serves it: the degree bound n is not divisible by the folding factor at the second layer
    from xfg_stark_b200 import air as A
    a5, t5 = A.power_map_air(2, 128, 5, seed=5); o = (25, 4, 2, 3, 16, 3)
    p = orc.prove_air(a5.flatten(), t5, o)
    assert orc.verify_air(p, a5.flatten(), o) == "DegreeTruncation" == goemul.verify_air(a5.flatten(), p, o)


def test_option_vectors_cover_the_option_space():
    cs = refvec.option_cases()
    assert {c["options"][3] for c in cs} == {1, 2, 3}
    assert {c["options"][1] for c in cs} >= {2, 4, 8, 16, 32, 64, 128}
    assert {c["options"][4] for c in cs} == {2, 4, 8, 16}
    assert {c["options"][5] for c in cs} >= {0, 1, 3, 7, 15, 31, 63, 255}
    assert {c["n_log2"] for c in cs} >= {6, 8, 9, 10, 11, 12, 13}


def test_shapes_the_reference_refuses_are_refused():
    ref = refvec.load_options()["refused"]
    assert len(ref) >= 3 and all("FRI layer" in r["panic"] or "TooFewLeaves" in r["panic"] for r in ref)
    for r in ref:
        o = tuple(r["options"]); n = 1 << r["n_log2"]
        assert refvec.fri_shape_refused(r["n_log2"], o)
        t, pi, ac = orc.synthetic_case(n, 1)
        with pytest.raises(goemul.EmulError) as e:
            goemul.prove_burn_mint(t, pi, ac, o)
        assert e.value.code == 2
    for c in refvec.option_cases() + refvec.cases():             # ... and nothing the reference proves is refused
        assert not refvec.fri_shape_refused(c["n_log2"], tuple(c["options"]))


def test_emulated_pipeline_equals_oracle_on_a_seeded_option_sweep():
    rng = random.Random(20261019)
    done = 0
    while done < 40:
        n_log2 = rng.choice([3, 4, 5, 6, 7, 8, 9])
        blowup = rng.choice([2, 4, 8, 16, 32, 64, 128]); folding = rng.choice([2, 4, 8, 16]); rem = rng.choice([0, 1, 3, 7, 15, 31, 63, 127, 255])
        o = (rng.randrange(1, min(255, (blowup << n_log2) - 1) + 1), blowup, rng.randrange(0, 9), rng.choice([1, 2, 3]), folding, rem)
        if refvec.fri_shape_refused(n_log2, o):
            continue
        index = rng.randrange(1 << 20)
        t, pi, ac = orc.synthetic_case(1 << n_log2, index)
        expect = orc.prove(t, pi, ac, o)
        assert goemul.prove_burn_mint(t, pi, ac, o) == expect, (n_log2, o)
        assert orc.verify(expect, pi, ac, o) == ""
        done += 1


def test_emulated_pipeline_detects_bad_traces():
    t, pi, ac = orc.synthetic_case(64, 3)
    o = (20, 4, 2, 3, 4, 3)
    bad = t.copy(); bad[4, 17] = 2
    with pytest.raises(goemul.EmulError) as e:
        goemul.prove_burn_mint(bad, pi, ac, o)
    assert e.value.code == 5
    bad = t.copy(); bad[2, 5] = orc.P + 1
    with pytest.raises(goemul.EmulError) as e:
        goemul.prove_burn_mint(bad, pi, ac, o)
    assert e.value.code == 1


ASAN_SWEEP = r"""
import random, sys
sys.path.insert(0, sys.argv[1])
import goemul, orc, refvec
rng = random.Random(7); done = 0
while done < 16:
    n_log2 = rng.choice([3, 4, 5, 6, 7, 8, 10, 11])
    blowup = rng.choice([2, 4, 8, 16, 32, 64, 128]); folding = rng.choice([2, 4, 8, 16]); rem = rng.choice([0, 1, 3, 7, 15, 31, 63, 127, 255])
    o = (rng.randrange(1, min(255, (blowup << n_log2) - 1) + 1), blowup, rng.randrange(0, 6), rng.choice([1, 2, 3]), folding, rem)
    if refvec.fri_shape_refused(n_log2, o):
        continue
    t, pi, ac = orc.synthetic_case(1 << n_log2, done)
    assert goemul.prove_burn_mint(t, pi, ac, o) == orc.prove(t, pi, ac, o), (n_log2, o)
    done += 1
print("asan sweep ok")
"""


def test_emulated_pipeline_under_address_sanitizer():
    """memory safety of the general-options kernel bodies: the index arithmetic the CUDA kernels run, executed under AddressSanitizer with a poisoned
    guard zone after every workspace region (compute-sanitizer is closed on the GPU pool)"""
    import os
    import subprocess
    import sys
    so, asan = goemul.build_asan()
    if so is None:
        pytest.skip("g++ has no libasan here")
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", GO_EMUL_LIB=so)
    r = subprocess.run([sys.executable, "-c", ASAN_SWEEP, here], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "asan sweep ok" in r.stdout, r.stderr[-2000:]
    r = subprocess.run([sys.executable, "-c", ASAN_SWEEP, here], env=dict(env, GO_EMUL_POKE="1"), capture_output=True, text=True, timeout=900)      # the checker itself
    assert r.returncode != 0 and "use-after-poison" in r.stderr
