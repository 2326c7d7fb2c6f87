"""Golden proofs produced by the REFERENCE's own prover (winterfell 0.8.3 inside /root/reference/test-dist/xfg-stark-cli, executed by
oracle/a64emu/make_reference_vectors.py) and the statement / trace / AIR each one belongs to.  Test infrastructure."""
import base64
import json
import os
import zlib

import numpy as np

import orc

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_proofs.json")


def load():
    return json.load(open(PATH))


def cases():
    return load()["cases"]


def case_ids():
    return [c["name"] for c in cases()]


def proof_bytes(c):
    return zlib.decompress(base64.b64decode(c["proof_zlib_b64"]))


def statement(c):
    """-> (pub_inputs, consts, options tuple, n)"""
    pi, ac, _ = orc.pack_inputs(8_000_000, 8_000_000, bytes.fromhex(c["tx_prefix_hash"]), bytes.fromhex(c["recipient"]), bytes.fromhex(c["secret"]),
                                c["network_id"], c["target_chain_id"], c["version"])
    return pi, ac, tuple(c["options"]), 1 << c["n_log2"]


def trace(c, pi, ac):
    """the reference's trace (src/burn_mint_air.rs:442-476): constants + state 0,1,2,3 over the quarters of the first 64 rows, then 3"""
    n = 1 << c["n_log2"]
    t = orc.build_trace(pi, ac, n)
    t[4] = np.minimum(3, np.arange(n) // 16)
    return t


def is_normalised(c):
    """True when the source's last-step assertion (step 63, src/burn_mint_air.rs:393) coincides with the normalised AIR's step n - 1"""
    return c["n_log2"] == 6


def air_program(c, pi, ac):
    """the burn-mint AIR with the source's literal assertion step, as a generic AIR description (for traces longer than 64 rows)"""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from xfg_stark_b200 import air as A
    return A.burn_mint_air(pi, ac[0], ac[1], ac[2], ac[3], 1 << c["n_log2"], last_step=c["last_step"])
