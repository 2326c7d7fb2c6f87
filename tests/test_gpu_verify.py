"""GPU parity tests of the batch verifier (xfg_verify_burn_mint_batch): its verdict on every proof - valid,
tampered byte by byte, truncated, proven for other public inputs or other options - must be the CPU oracle verifier's verdict.

The oracle names its parse-level rejections individually ("truncated proof", "bad OOD frame", ...); the C ABI folds them into one
code (ProofDeserializationError).  Cryptographic rejections carry winterfell's VerifierError names on both sides and must match
exactly, with one documented difference: the ABI checks section lengths and the canonicity of the OOD frame / remainder
before the transcript replay, so a proof that is both malformed there and cryptographically wrong is reported as malformed."""
import numpy as np
import pytest

import orc
from test_gpu_proof import gpu_case

pytestmark = pytest.mark.gpu

CRYPTO = {"UnacceptableProofOptions", "InconsistentOodConstraintEvaluations", "QuerySeedProofOfWorkVerificationFailed", "NumberOfQueriesMismatch",
          "TraceQueryDoesNotMatchCommitment", "ConstraintQueryDoesNotMatchCommitment", "LayerCommitmentMismatch", "InvalidLayerFolding",
          "RemainderCommitmentMismatch", "RemainderDegreeMismatch", "InvalidRemainderFolding", "DegreeTruncation"}


def same_verdict(ours, oracle):
    if oracle == "":
        return ours == ""
    if ours == "":
        return False
    if ours == "ProofDeserializationError":      # see the module docstring
        return True
    return (oracle in CRYPTO and ours == oracle)


def tamper_offsets(proof, rng, count):
    """byte offsets spread over the whole proof: the first 64 bytes (context, options, commitments), the tail (remainder, nonce),
    and random positions in between (query values, Merkle paths, OOD frame, FRI layers)"""
    n = len(proof)
    offs = set(range(0, min(64, n))) | set(range(max(0, n - 48), n)) | {int(x) for x in rng.integers(0, n, size=count)}
    return sorted(offs)


@pytest.mark.parametrize("ext", [1, 2])
@pytest.mark.parametrize("n_log2", [3, 6, 10, 13])
def test_valid_proofs_accepted(ctx, n_log2, ext):
    import xfg_stark_b200 as xs
    opts = xs.ProofOptions(field_extension=ext)
    air, trace = gpu_case(xs, n_log2, n_log2)
    proof = ctx.prove(trace, air, opts)
    tr, pi, ac = orc.synthetic_case(1 << n_log2, n_log2)
    assert orc.verify(proof, pi, ac, opts.as_tuple()) == ""
    assert ctx.verify_batch([proof], [air], opts) == [""]
    v = xs.XfgBurnMintVerifier(128, opts, context=ctx)
    assert v.verify_with_public_inputs(proof, air) is True
    v.verify_with_winterfell(proof, air)


@pytest.mark.parametrize("opts_t", [(42, 8, 0, 1, 8, 31), (1, 8, 4, 2, 8, 7), (255, 8, 10, 1, 8, 255), (27, 8, 16, 2, 8, 63), (100, 8, 20, 1, 8, 15)])
def test_option_sweep_accepted_and_cross_rejected(ctx, opts_t):
    import xfg_stark_b200 as xs
    opts = xs.ProofOptions(*opts_t)
    air, trace = gpu_case(xs, 11, 11)
    proof = ctx.prove(trace, air, opts)
    assert ctx.verify_batch([proof], [air], opts) == [""]
    other = xs.ProofOptions()                       # the default option set does not accept this proof
    tr, pi, ac = orc.synthetic_case(1 << 11, 11)
    assert orc.verify(proof, pi, ac, other.as_tuple()) == "UnacceptableProofOptions"
    assert ctx.verify_batch([proof], [air], other) == ["UnacceptableProofOptions"]


@pytest.mark.parametrize("n_log2,ext", [(6, 1), (6, 2), (10, 2), (12, 1)])
def test_every_tampered_byte_gets_the_oracle_verdict(ctx, n_log2, ext):
    import xfg_stark_b200 as xs
    opts = xs.ProofOptions(field_extension=ext)
    air, trace = gpu_case(xs, 3, n_log2)
    proof = ctx.prove(trace, air, opts)
    tr, pi, ac = orc.synthetic_case(1 << n_log2, 3)
    rng = np.random.default_rng(n_log2 * 10 + ext)
    offs = tamper_offsets(proof, rng, 300)
    bad = []
    for o in offs:
        b = bytearray(proof); b[o] ^= 1 << int(rng.integers(0, 8)); bad.append(bytes(b))
    bad.append(proof[:-1]); bad.append(proof + b"\0"); bad.append(proof[:len(proof) // 2]); bad.append(b""); bad.append(proof)
    ours = ctx.verify_batch(bad, [air] * len(bad), opts)
    theirs = [orc.verify(p, pi, ac, opts.as_tuple()) for p in bad]
    assert ours[-1] == "" and theirs[-1] == ""
    mism = [(i, a, b) for i, (a, b) in enumerate(zip(ours, theirs)) if not same_verdict(a, b)]
    assert not mism, mism[:10]
    assert sum(1 for a in ours[:-1] if a == "") == 0           # every single-bit change is rejected
    # the comparison is not vacuous: several distinct cryptographic checks fired, with identical names on both sides
    fired = {a for a, b in zip(ours, theirs) if a in CRYPTO and a == b}
    assert len(fired) >= 4, fired


def test_wrong_statement_rejected(ctx):
    """a valid proof checked against other public inputs / AIR constants"""
    import xfg_stark_b200 as xs
    opts = xs.ProofOptions()
    air0, trace0 = gpu_case(xs, 0, 8)
    air1, _ = gpu_case(xs, 1, 8)
    proof = ctx.prove(trace0, air0, opts)
    _, pi1, ac1 = orc.synthetic_case(256, 1)
    ours = ctx.verify_batch([proof, proof], [air0, air1], opts)
    assert ours[0] == "" and ours[1] != ""
    assert same_verdict(ours[1], orc.verify(proof, pi1, ac1, opts.as_tuple()))


def test_mixed_batch_and_mirror_classes(ctx):
    """BatchBurnMintVerifier::verify_batch / verify_all (src/burn_mint_verifier.rs:371-408): 64 proofs of different statements and
    lengths in one launch, a few of them corrupted"""
    import xfg_stark_b200 as xs
    opts = xs.ProofOptions()
    items, expect = [], []
    rng = np.random.default_rng(5)
    for i in range(64):
        n_log2 = 6 + i % 5
        air, trace = gpu_case(xs, i, n_log2)
        proof = ctx.prove(trace, air, opts)
        ok = i % 7 != 3
        if not ok:
            b = bytearray(proof); b[int(rng.integers(100, len(proof) - 60))] ^= 0x10; proof = bytes(b)
        items.append((proof, air)); expect.append(ok)
    bv = xs.BatchBurnMintVerifier(128, opts, context=ctx)
    assert bv.verify_batch(items) == expect
    assert bv.verify_all(items) is False
    assert bv.verify_all([it for it, ok in zip(items, expect) if ok]) is True
    with pytest.raises(xs.XfgError) as e:
        bv.verifier.verify_with_winterfell(*items[3])
    assert "STARK verification failed" in str(e.value)


def test_oracle_proofs_and_large_proof_accepted():
    """proofs produced by the CPU oracle are accepted too (the verifier shares no state with the prover), incl. a 2^16-row one"""
    import xfg_stark_b200 as xs
    from test_gpu_stages import big_ctx
    c = big_ctx()
    for n_log2, ext in [(9, 1), (16, 2)]:
        opts = xs.ProofOptions(field_extension=ext)
        air, _ = gpu_case(xs, 2, n_log2)
        tr, pi, ac = orc.synthetic_case(1 << n_log2, 2)
        orc.set_threads(orc.max_threads())
        proof = orc.prove(tr, pi, ac, opts.as_tuple())
        orc.set_threads(1)
        res, times = c.verify_batch([proof] * 8, [air] * 8, opts, want_times=True)
        assert res == [""] * 8 and times["kernel_ms"] > 0 and times["h2d_bytes"] >= 8 * len(proof)
