#!/usr/bin/env python
"""make_reference_option_vectors.py — golden proofs from the REFERENCE ITSELF for `ProofOptions` other than the reference's default
(`XfgBurnMintProver::with_options`, src/burn_mint_prover.rs:44-49): blowup factors 2 .. 128, FRI folding factors 2 / 4 / 8 / 16, remainder
degrees 0 .. 255 and all three field extensions (None, Quadratic, Cubic).  Same method and the same two interventions as
make_reference_vectors.py (the Winterfell 0.8.3 prover linked into /root/reference/test-dist/xfg-stark-cli, executed by the a64emu
interpreter); writes tests/golden/reference_proofs_options.json.  Also records the option sets the reference itself refuses with a panic.

Needs /root/reference (this container only).  The vectors travel as a committed fixture.
"""
import base64
import hashlib
import json
import os
import sys
import time
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import refbin  # noqa: E402
from make_reference_vectors import Reference, long_trace  # noqa: E402

# (num_queries, blowup, grinding, extension, folding, remainder_max_degree)
SHORT = [(42, 2, 4, 1, 2, 0), (42, 4, 4, 2, 4, 3), (42, 16, 4, 2, 16, 7), (20, 32, 0, 1, 2, 1), (10, 128, 2, 1, 4, 255), (42, 8, 4, 2, 8, 3),
         (42, 8, 4, 1, 4, 1), (42, 8, 4, 3, 8, 31), (42, 2, 4, 3, 2, 0), (42, 4, 4, 3, 4, 3), (42, 16, 4, 3, 16, 7), (20, 32, 0, 3, 2, 1),
         (33, 64, 5, 2, 16, 15), (42, 8, 4, 2, 2, 31), (42, 8, 4, 1, 16, 31), (42, 4, 4, 1, 8, 31), (42, 16, 4, 2, 8, 31)]
LONG = [(8, (42, 8, 4, 3, 8, 31)), (9, (42, 4, 4, 2, 2, 7)), (10, (42, 16, 4, 1, 4, 15)), (11, (30, 2, 3, 3, 16, 7)), (12, (42, 32, 2, 2, 8, 63)),
        (12, (42, 8, 4, 3, 8, 31)), (13, (27, 4, 8, 3, 4, 1))]
# option sets ProofOptions::new accepts but the prover cannot serve at 64 rows: recorded with the reference's own panic message
REFUSED = [(6, (30, 4, 0, 2, 16, 0)), (11, (30, 2, 3, 3, 16, 3)), (9, (27, 4, 8, 3, 4, 0))]   # a FRI layer of one row (x2); an empty remainder


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    ref = Reference()
    out = {"source": "winterfell 0.8.3 as linked into /root/reference/test-dist/xfg-stark-cli, executed by oracle/a64emu (see make_reference_option_vectors.py)",
           "binary_sha256": hashlib.sha256(ref.rb.m.data).hexdigest(), "cases": [], "refused": []}

    def inputs(k):
        g = orc.splitmix64(0x4F5054494F4E + k)
        raw = b"".join(next(g).to_bytes(8, "little") for _ in range(11))
        return raw[:32], raw[32:52], bytes([1, 2, 3, 4]) + raw[60:88]

    def add(name, n_log2, options, k, long_):
        tx, rcpt, secret = inputs(k)
        pi, ac, _ = orc.pack_inputs(8_000_000, 8_000_000, tx, rcpt, secret, 4, 42161, 1)
        t0 = time.time()
        if long_:
            proof, ic = ref.prove_long(tx, rcpt, secret, options, long_trace(pi, ac, 1 << n_log2))
        else:
            proof, ic = ref.prove64(tx, rcpt, secret, options)
        print(f"{name}: {len(proof)} bytes, {ic / 1e6:.1f} M guest instructions, {time.time() - t0:.1f} s", flush=True)
        out["cases"].append({"name": name, "n_log2": n_log2, "options": list(options), "last_step": 63, "tx_prefix_hash": tx.hex(), "recipient": rcpt.hex(),
                             "secret": secret.hex(), "network_id": 4, "target_chain_id": 42161, "version": 1, "entry": "Prover::prove" if long_ else "prove_burn_mint",
                             "guest_instructions": ic, "proof_sha256": hashlib.sha256(proof).hexdigest(), "proof_len": len(proof),
                             "proof_zlib_b64": base64.b64encode(zlib.compress(proof, 9)).decode()})

    def tag(o):
        return "q%d_b%d_g%d_e%d_f%d_r%d" % o

    k = 0
    for o in SHORT:
        add("n64_" + tag(o), 6, o, k, False); k += 1
    for n_log2, o in LONG:
        add(f"n2p{n_log2}_" + tag(o), n_log2, o, k, True); k += 1
    for n_log2, o in REFUSED:
        fresh = Reference()
        tx, rcpt, secret = inputs(k); k += 1
        pi, ac, _ = orc.pack_inputs(8_000_000, 8_000_000, tx, rcpt, secret, 4, 42161, 1)
        try:
            if n_log2 == 6:
                fresh.prove64(tx, rcpt, secret, o)
            else:
                fresh.prove_long(tx, rcpt, secret, o, long_trace(pi, ac, 1 << n_log2))
            msg = "returned a proof (unexpected)"
        except refbin.EmuError:
            txt = fresh.rb.text_output()
            msg = " ".join(txt[txt.find("panicked at"):].split("\n")[1:2]).strip()
        print("refused", n_log2, o, "->", msg)
        out["refused"].append({"n_log2": n_log2, "options": list(o), "panic": msg})
    path = os.path.join(ROOT, "tests", "golden", "reference_proofs_options.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
