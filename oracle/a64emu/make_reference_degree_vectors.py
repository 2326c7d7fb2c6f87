#!/usr/bin/env python
"""make_reference_degree_vectors.py — golden proofs from the REFERENCE ITSELF for AIRs whose transition constraints DECLARE degrees above 2
(`TransitionConstraintDegree::new(d)`, the shape of src/burn_mint_air.rs:103-113 with another number): winter-air then derives d - 1 constraint
composition columns and a constraint-evaluation blowup of next_pow2(d - 1).  Same method as make_reference_vectors.py (the Winterfell 0.8.3 prover linked
into /root/reference/test-dist/xfg-stark-cli, executed by the a64emu interpreter); the AirContext intervention passes [2, d, 1, 1, 2, 1, 1] as the seven
declared degrees.  The burn-mint constraints themselves stay what they are, so the upper columns are zero: these vectors pin the LAYOUT of a
multi-column proof (column count, evaluation domain, OOD evaluations, DEEP coefficients, query rows); real degree-d constraints are covered by the
oracle's prover / verifier pair and the product against the oracle (tests/test_options_pins.py, tests/test_gpu_options.py).
Writes tests/golden/reference_proofs_degrees.json.  Needs /root/reference (this container only)."""
import base64
import hashlib
import json
import os
import struct
import sys
import time
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_reference_vectors import Reference, long_trace  # noqa: E402


class DegreeReference(Reference):
    degs = [1] * 7

    def _fix_ctx(self, r):
        cap, ptr, ln = r.u64s(r.x[1], 3)
        first = list(struct.unpack("<4Q", r.read(ptr, 32)))      # TransitionConstraintDegree { cycles: Vec<usize> (cap, ptr, len), base: usize }
        at = first.index(1)
        blob = b""
        for d in self.degs:
            w = list(first); w[at] = d
            blob += struct.pack("<4Q", *w)
        r.write(r.x[1], struct.pack("<QQQ", 7, r.put(blob), 7))
        r.x[2] = 8
        return "continue"


CASES = [(6, 3, (42, 8, 4, 1, 8, 31)), (6, 3, (42, 8, 4, 2, 8, 31)), (6, 3, (30, 16, 2, 3, 4, 7)), (6, 4, (42, 8, 4, 1, 8, 31)), (6, 4, (42, 8, 4, 2, 8, 31)),
         (6, 4, (30, 4, 2, 3, 4, 7)), (6, 5, (42, 8, 4, 2, 8, 31)), (6, 5, (20, 32, 0, 1, 2, 1)), (6, 6, (42, 8, 4, 2, 8, 31)), (6, 9, (42, 16, 4, 3, 16, 7)),
         (9, 3, (42, 8, 4, 2, 8, 31)), (10, 4, (42, 4, 4, 1, 4, 15)), (11, 5, (30, 8, 3, 3, 8, 31))]


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    out = {"source": "winterfell 0.8.3 as linked into /root/reference/test-dist/xfg-stark-cli, executed by oracle/a64emu (see make_reference_degree_vectors.py)", "cases": []}
    for k, (n_log2, d, options) in enumerate(CASES):
        ref = DegreeReference(); ref.degs = [2, d, 1, 1, 2, 1, 1]
        g = orc.splitmix64(0x444547524545 + k)
        raw = b"".join(next(g).to_bytes(8, "little") for _ in range(11))
        tx, rcpt, secret = raw[:32], raw[32:52], bytes([1, 2, 3, 4]) + raw[60:88]
        pi, ac, _ = orc.pack_inputs(8_000_000, 8_000_000, tx, rcpt, secret, 4, 42161, 1)
        t0 = time.time()
        if n_log2 == 6:
            proof, ic = ref.prove64(tx, rcpt, secret, options)
        else:
            proof, ic = ref.prove_long(tx, rcpt, secret, options, long_trace(pi, ac, 1 << n_log2))
        name = "n2p%d_deg%d_q%d_b%d_g%d_e%d_f%d_r%d" % ((n_log2, d) + tuple(options))
        print(f"{name}: {len(proof)} bytes, {ic / 1e6:.1f} M guest instructions, {time.time() - t0:.1f} s", flush=True)
        out["binary_sha256"] = hashlib.sha256(ref.rb.m.data).hexdigest()
        out["cases"].append({"name": name, "n_log2": n_log2, "declared_degree": d, "options": list(options), "last_step": 63, "tx_prefix_hash": tx.hex(), "recipient": rcpt.hex(),
                             "secret": secret.hex(), "network_id": 4, "target_chain_id": 42161, "version": 1, "entry": "Prover::prove" if n_log2 != 6 else "prove_burn_mint",
                             "guest_instructions": ic, "proof_sha256": hashlib.sha256(proof).hexdigest(), "proof_len": len(proof),
                             "proof_zlib_b64": base64.b64encode(zlib.compress(proof, 9)).decode()})
    path = os.path.join(ROOT, "tests", "golden", "reference_proofs_degrees.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
