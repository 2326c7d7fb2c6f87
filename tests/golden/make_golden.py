"""Regenerates tests/golden/proofs.json from the CPU oracle:  python tests/golden/make_golden.py

These are REGRESSION pins of this repo's own oracle (the reference holds no golden vectors for the proving path and its
Winterfell crates cannot be built here - SURVEY.md §8c), so that any later change of oracle or CUDA output is caught; they are
not outputs of the reference."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import orc  # noqa: E402

out = {"note": "sha256 of oracle proof bytes for synthetic_case(n, index=log2 n); options (42,8,4,ext,8,31)", "cases": []}
for n_log2 in (3, 6, 10, 12):
    for ext in (1, 2):
        tr, pi, ac = orc.synthetic_case(1 << n_log2, n_log2)
        proof = orc.prove(tr, pi, ac, (42, 8, 4, ext, 8, 31), keep_debug=True)
        out["cases"].append({"n_log2": n_log2, "ext": ext, "len": len(proof), "sha256": hashlib.sha256(proof).hexdigest(),
                             "trace_root": bytes(orc.debug_get("trace_root").view("u1")).hex(), "nonce": int(orc.debug_get("nonce")[0]),
                             "num_positions": int(orc.debug_get("positions").size),
                             "proof_hex": proof.hex() if n_log2 == 3 and ext == 1 else None})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "proofs.json"), "w"), indent=1)
print("wrote", len(out["cases"]), "cases")

# ---- generic AIR front-end: regression pins of the oracle's generic path on the example AIRs of tests/test_air.py
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import test_air  # noqa: E402

air_out = {"note": "sha256 of oracle proof bytes of the example AIRs of tests/test_air.py::examples(); options (42,8,4,ext,8,31)"}
for name, (air, trace) in test_air.examples().items():
    for ext in (1, 2):
        air_out[f"{name}/ext{ext}"] = hashlib.sha256(orc.prove_air(air.flatten(), trace, (42, 8, 4, ext, 8, 31))).hexdigest()
json.dump(air_out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "air_proofs.json"), "w"), indent=1)
print("wrote", len(air_out) - 1, "generic-AIR cases")
