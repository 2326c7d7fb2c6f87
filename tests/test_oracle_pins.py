"""CPU tests (-m "not gpu"): pin the oracle (oracle/, test infrastructure) against every independent source available here:
Python `blake3`, the reference's Keccak KAT (src/lib.rs:141-148), Python big integers for the field / extension / NTT algebra,
the wire-format examples of SURVEY.md Appendix A, and the committed regression fixture tests/golden/proofs.json.
Proof-byte parity with real Winterfell 0.8.3 is pinned separately, against proofs emitted by the reference's own binary
(tests/test_reference_binary_pins.py, tests/golden/reference_proofs.json)."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import orc

P = orc.P
HERE = os.path.dirname(os.path.abspath(__file__))


def test_blake3_matches_python_blake3():
    import blake3
    rng = random.Random(1)
    for ln in [0, 1, 8, 16, 40, 56, 63, 64, 65, 112, 128, 160, 224, 512, 1023, 1024, 1025, 2048, 2049, 3072, 4096, 5000, 9000]:
        d = bytes(rng.getrandbits(8) for _ in range(ln))
        assert orc.blake3(d) == blake3.blake3(d).digest(), ln


def test_keccak256_reference_kat():
    # src/lib.rs:141-148: the only known-answer test in the reference
    assert orc.keccak256(b"93385046440755750514194170694064996624").hex() == "6430829be74c2d9892a5122aa2f2daac3ee9850f086a8985941e7fb4bde60fcf"
    h = orc.keccak256(b"93385046440755750514194170694064996624")
    assert int.from_bytes(h[:8], "little") % ((1 << 63) - 1) == 1742133188492406885      # src/lib.rs:150-160
    assert orc.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert orc.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    # (every Keccak message on the proving path is a single 136-byte block: the longest, the commitment preimage, is 130 bytes)


def test_field_arithmetic_against_bigint():
    rng = random.Random(2)
    edge = [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 1 << 32, (1 << 32) + 1, 1 << 63, 0xFFFFFFFF00000000, P - (1 << 32)]
    vals = edge + [rng.randrange(P) for _ in range(300)]
    L = orc.lib()
    for a in vals:
        for b in vals[:40]:
            assert L.orc_fmul(a, b) == a * b % P == L.orc_fmul_slow(a, b)
            assert L.orc_fadd(a, b) == (a + b) % P and L.orc_fsub(a, b) == (a - b) % P
        if a:
            assert L.orc_fmul(a, L.orc_finv(a)) == 1


def test_roots_of_unity_pinned_by_the_reference_binary():
    # 2^32-th root 7277203076849721926 (the binary's constant), hence w_8 = 2^24 and w_64 = 8
    g = 7277203076849721926
    assert orc.root_of_unity(32) == g and pow(g, 1 << 32, P) == 1 and pow(g, 1 << 31, P) == P - 1
    assert orc.root_of_unity(3) == 1 << 24 and orc.root_of_unity(6) == 8 and orc.root_of_unity(1) == P - 1
    assert pow(7, (P - 1) // 2, P) == P - 1          # 7 generates the multiplicative group's 2-part: a valid coset offset


def f2mul(a, b):   # F_p[x]/(x^2 - x + 2), schoolbook: x^2 = x - 2
    c0 = a[0] * b[0]; c1 = a[0] * b[1] + a[1] * b[0]; c2 = a[1] * b[1]
    return ((c0 - 2 * c2) % P, (c1 + c2) % P)


def test_quadratic_extension_against_schoolbook():
    rng = random.Random(3); L = orc.lib()
    for _ in range(200):
        a = np.array([rng.randrange(P), rng.randrange(P)], dtype=np.uint64); b = np.array([rng.randrange(P), rng.randrange(P)], dtype=np.uint64)
        o = np.zeros(2, dtype=np.uint64)
        L.orc_f2_mul(orc._p(a), orc._p(b), orc._p(o))
        assert (int(o[0]), int(o[1])) == f2mul((int(a[0]), int(a[1])), (int(b[0]), int(b[1])))
        L.orc_f2_inv(orc._p(a), orc._p(o))
        assert f2mul((int(a[0]), int(a[1])), (int(o[0]), int(o[1]))) == (1, 0)


@pytest.mark.parametrize("n_log2", [1, 3, 6, 8])
@pytest.mark.parametrize("deg", [1, 2])
def test_ntt_matches_naive_dft_and_inverts(n_log2, deg):
    rng = np.random.default_rng(n_log2 + deg)
    n = 1 << n_log2
    a = (rng.integers(0, 1 << 62, size=n * deg, dtype=np.uint64) % np.uint64(P))
    f = orc.ntt(a, deg, 0)
    assert (f == orc.ntt(a, deg, 2)).all()
    assert (orc.ntt(f, deg, 1) == a).all()


def test_lde_is_evaluation_on_the_coset():
    """evaluate_poly_with_offset(p, 7, 8)[i] == p(7 * w_N^i) by Horner with big integers (A.7)"""
    rng = random.Random(5); n = 16; N = 8 * n
    coef = [rng.randrange(P) for _ in range(n)]
    lde = orc.lde(np.array(coef, dtype=np.uint64))
    wN = orc.root_of_unity(7)
    for i in range(N):
        x = 7 * pow(wN, i, P) % P; acc = 0
        for c in reversed(coef):
            acc = (acc * x + c) % P
        assert int(lde[i]) == acc
    back = orc.interpolate_offset(lde)
    assert [int(v) for v in back[:n]] == coef and not back[n:].any()


def test_merkle_tree_layout_and_batch_proof():
    import blake3
    rng = np.random.default_rng(6)
    leaves = rng.integers(0, 256, size=(8, 32), dtype=np.uint8)
    root, nodes = orc.merkle(leaves)
    h = lambda a, b: blake3.blake3(bytes(a) + bytes(b)).digest()
    n4 = [h(leaves[2 * i], leaves[2 * i + 1]) for i in range(4)]
    n2 = [h(n4[0], n4[1]), h(n4[2], n4[3])]
    assert root == h(n2[0], n2[1])
    assert bytes(nodes[4]) == n4[0] and bytes(nodes[7]) == n4[3] and bytes(nodes[2]) == n2[0] and bytes(nodes[1]) == root   # A.7 numbering
    # A.11: prove_batch([1, 6]) -> two node vectors: [leaf0, n4[1]] and [leaf7, n4[2]]
    pf = orc.merkle_prove_batch(leaves, [1, 6])
    assert pf == bytes([2, 2]) + bytes(leaves[0]) + n4[1] + bytes([2]) + bytes(leaves[7]) + n4[2]
    # both leaves of a pair queried: nothing at leaf level, sibling pair node only
    pf = orc.merkle_prove_batch(leaves, [2, 3])
    assert pf == bytes([1, 2]) + n4[0] + n2[1]


def test_wire_format_examples_from_the_survey():
    tr, pi, ac = orc.synthetic_case(64, 0)
    proof = orc.prove(tr, pi, ac)
    assert proof[:21].hex() == "07000006000008010000" "00ffffffff2a080401081f"          # A.12 Context example
    q = orc.prove(tr, pi, ac, (42, 8, 4, 2, 8, 31))
    assert q[:21].hex() == "07000006000008010000" "00ffffffff2a080402081f"
    assert proof[22:24] == (32 * 4).to_bytes(2, "little")                               # commitments: trace, constraint, 1 FRI layer, remainder
    assert proof[-8:] == int(orc_nonce(tr, pi, ac)).to_bytes(8, "little") and proof[-9] == 0   # ... log2(num_partitions) = 0, pow_nonce


def orc_nonce(tr, pi, ac, opts=orc.DEFAULT_OPTIONS):
    orc.prove(tr, pi, ac, opts, keep_debug=True)
    return orc.debug_get("nonce")[0]


def test_grinding_nonce_is_the_smallest():
    tr, pi, ac = orc.synthetic_case(64, 1)
    for g in (0, 4, 9):
        opts = (42, 8, g, 1, 8, 31)
        nonce = int(orc_nonce(tr, pi, ac, opts))
        assert nonce >= 1 and (g > 0 or nonce == 1)


@pytest.mark.parametrize("ext", [1, 2])
@pytest.mark.parametrize("n_log2", [3, 5, 6, 9])
def test_oracle_verifier_accepts_and_rejects(n_log2, ext):
    tr, pi, ac = orc.synthetic_case(1 << n_log2, 7)
    opts = (42, 8, 4, ext, 8, 31)
    proof = orc.prove(tr, pi, ac, opts)
    assert orc.verify(proof, pi, ac, opts) == ""
    rng = random.Random(n_log2 * 2 + ext)
    for _ in range(60):                                    # any single-bit tamper is rejected
        b = bytearray(proof); i = rng.randrange(len(b)); b[i] ^= 1 << rng.randrange(8)
        assert orc.verify(bytes(b), pi, ac, opts) != "", i
    pi2 = pi.copy(); pi2[9] += 1                           # other network id: different coin seed
    assert orc.verify(proof, pi2, ac, opts) != ""
    assert orc.verify(proof, pi, ac, (41, 8, 4, ext, 8, 31)) == "UnacceptableProofOptions"
    assert orc.verify(proof[:-1], pi, ac, opts) != ""


def test_oracle_is_deterministic_and_thread_invariant():
    tr, pi, ac = orc.synthetic_case(1 << 12, 3)
    a = orc.prove(tr, pi, ac)
    orc.set_threads(orc.max_threads())
    b = orc.prove(tr, pi, ac)
    orc.set_threads(1)
    assert a == b == orc.prove(tr, pi, ac)


def test_unsatisfied_trace_is_refused():
    tr, pi, ac = orc.synthetic_case(64, 0)
    bad = tr.copy(); bad[4, 20] = 3
    with pytest.raises(RuntimeError, match="UnsatisfiedTransitionConstraintError"):
        orc.prove(bad, pi, ac)


def test_golden_regression_fixture():
    cases = json.load(open(os.path.join(HERE, "golden", "proofs.json")))["cases"]
    for c in cases:
        tr, pi, ac = orc.synthetic_case(1 << c["n_log2"], c["n_log2"])
        proof = orc.prove(tr, pi, ac, (42, 8, 4, c["ext"], 8, 31))
        assert len(proof) == c["len"] and hashlib.sha256(proof).hexdigest() == c["sha256"], c
        if c["proof_hex"]:
            assert proof.hex() == c["proof_hex"]


def test_input_packing_follows_the_reference():
    """src/burn_mint_prover.rs:62-107, 132-221 and src/burn_mint_air.rs:124-202"""
    s = orc.synthetic_inputs(0)
    pi, ac, se = orc.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], 4, 42161, 1)
    assert se == int.from_bytes(s["secret"][:4], "little")
    assert int(pi[0]) == int(pi[1]) == 8_000_000 and int(pi[4]) == 0
    assert int(pi[2]) == int.from_bytes(s["tx_prefix_hash"][:4], "little") == int(pi[5])
    assert [int(pi[5 + i]) for i in range(4)] == [int.from_bytes(s["tx_prefix_hash"][4 * i:4 * i + 4], "little") for i in range(4)]
    assert int(pi[3]) == int.from_bytes(orc.keccak256(s["recipient"] + b"recipient")[:4], "little")
    le = lambda v: int(v).to_bytes(8, "little")
    assert int(ac[2]) == int.from_bytes(orc.keccak256(le(se) + b"nullifier" + le(pi[0]))[:4], "little")
    rfull = orc.keccak256(le(pi[3]) + b"ethereum-recipient" + b"fuego-to-heat-bridge")
    pre = le(se) + le(pi[0]) + le(pi[1]) + b"".join(le(pi[5 + i]) for i in range(4)) + rfull + le(pi[9]) + le(pi[10]) + le(pi[11]) + b"heat-commitment-v1"
    assert int(ac[3]) == int.from_bytes(orc.keccak256(pre)[:4], "little")
    for bad, msg in [(dict(burn=1000, mint=1000), "Burn amount must be exactly"), (dict(mint=8_000_000_000), "does not match burn amount"),
                     (dict(tx_prefix_hash=bytes(32)), "Transaction hash must be greater than 0"), (dict(recipient=b"x" * 19), "exactly 20 bytes"),
                     (dict(secret=b"abc"), "at least 4 bytes")]:
        k = dict(s); k.update(bad)
        with pytest.raises(ValueError, match=msg):
            orc.pack_inputs(k["burn"], k["mint"], k["tx_prefix_hash"], k["recipient"], k["secret"], 4, 42161, 1)
    # defect B.1-4: the 800 XFG tier passes validation but truncates to u32 and cannot satisfy constraint 0
    pi8, ac8, _ = orc.pack_inputs(8_000_000_000, 8_000_000_000, s["tx_prefix_hash"], s["recipient"], s["secret"], 4, 42161, 1)
    assert int(pi8[0]) == 8_000_000_000 % (1 << 32)
    with pytest.raises(RuntimeError):
        orc.prove(orc.build_trace(pi8, ac8, 64), pi8, ac8)
