"""GPU parity against the REFERENCE's own proofs: bytes emitted by the CUDA backend (through the C ABI) must equal the bytes the
reference's Winterfell 0.8.3 prover emitted for the same statement, trace and options (tests/golden/reference_proofs.json, produced by
executing /root/reference/test-dist/xfg-stark-cli under oracle/a64emu - see oracle/a64emu/make_reference_vectors.py)."""
import numpy as np
import pytest

import orc
import refvec

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", refvec.case_ids())
def test_gpu_proof_equals_reference_proof(name):
    import xfg_stark_b200 as xs
    c = next(x for x in refvec.cases() if x["name"] == name)
    ref = refvec.proof_bytes(c)
    pi, ac, o, n = refvec.statement(c)
    opts = xs.ProofOptions(num_queries=o[0], blowup_factor=o[1], grinding_factor=o[2], field_extension=o[3], fri_folding_factor=o[4], fri_remainder_max_degree=o[5])
    t = refvec.trace(c, pi, ac)
    with xs.Context(device=0, max_n_log2=c["n_log2"]) as ctx:
        # generic AIR front-end with the source's literal assertion step (src/burn_mint_air.rs:393): any trace length
        assert ctx.prove_air(refvec.air_program(c, pi, ac), t, opts) == ref
        if refvec.is_normalised(c):
            # 64 rows: the hand-written burn-mint kernels, through every entry point of the reference's path
            air = xs.pack_inputs(8_000_000, 8_000_000, bytes.fromhex(c["tx_prefix_hash"]), bytes.fromhex(c["recipient"]), bytes.fromhex(c["secret"]),
                                 c["network_id"], c["target_chain_id"], c["version"])
            assert ctx.prove(t, air, opts) == ref
            assert ctx.prove_from_inputs(8_000_000, 8_000_000, bytes.fromhex(c["tx_prefix_hash"]), bytes.fromhex(c["recipient"]), bytes.fromhex(c["secret"]),
                                         c["network_id"], c["target_chain_id"], c["version"], n_log2=6, options=opts) == ref
            R = (1 << 64) % orc.P
            mont = [np.array([(int(v) * R) % orc.P for v in t[k]], dtype=np.uint64) for k in range(7)]     # TraceTable memory: Montgomery form
            assert ctx.prove_cols(mont, air, opts, form=1) == ref
            assert ctx.verify_batch([ref], [air], opts) == [""]                                           # the CUDA verifier accepts the reference's proof
