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


# proofs for `ProofOptions` other than the reference's default (blowup 2..128, folding 2/4/8/16, remainder degree 0..255, all three extensions):
# oracle/a64emu/make_reference_option_vectors.py
OPTIONS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_proofs_options.json")


def load_options():
    return json.load(open(OPTIONS_PATH))


def option_cases():
    return load_options()["cases"]


def option_case_ids():
    return [c["name"] for c in option_cases()]


# proofs of AIRs that DECLARE transition degrees above 2 (several constraint composition columns): oracle/a64emu/make_reference_degree_vectors.py
DEGREES_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_proofs_degrees.json")


def degree_cases():
    return json.load(open(DEGREES_PATH))["cases"]


def degree_case_ids():
    return [c["name"] for c in degree_cases()]


def fri_shape_refused(n_log2, options):
    """the rule behind the reference's own panics (fixture key "refused"): a FRI layer of fewer than two rows, or an empty remainder"""
    _, blowup, _, _, folding, rem = options
    lb, lf = blowup.bit_length() - 1, folding.bit_length() - 1
    l = n_log2 + lb
    while (1 << l) > (rem + 1) * blowup:
        if l < lf + 1:
            return True
        l -= lf
    return l < lb


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
    return A.burn_mint_air(pi, ac[0], ac[1], ac[2], ac[3], 1 << c["n_log2"], last_step=c["last_step"], pad_degree=c.get("declared_degree"))
