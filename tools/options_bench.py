#!/usr/bin/env python
"""options_bench.py — device time of the general-options pipeline (`with_options`, src/burn_mint_prover.rs:44-49) for a few option sets, next to
the tuned 8/8 pipeline on the same trace.  Every proof is checked against the CPU oracle when --check is given (small sizes).  One JSON line.
Times: CUDA events on the proof's stream (xfg_stage_times.device_ms), median of `--reps` proofs after one warm-up, trace resident in HBM."""
import argparse
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-log2", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch
    import xfg_stark_b200 as xs
    import orc
    n_log2 = args.n_log2
    s = orc.synthetic_inputs(0)
    air = xs.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
    trace = xs.build_trace(air, n_log2)
    sets = [("tuned_default_quadratic", (42, 8, 4, 2, 8, 31)), ("cubic_b8_f8", (42, 8, 4, 3, 8, 31)), ("quadratic_b8_f4", (42, 8, 4, 2, 4, 31)), ("quadratic_b16_f8", (42, 16, 4, 2, 8, 31)),
            ("quadratic_b4_f2_rem7", (42, 4, 4, 2, 2, 7)), ("none_b2_f16", (42, 2, 4, 1, 16, 31)), ("cubic_b32_f4", (42, 32, 4, 3, 4, 15))]
    rows = []
    with xs.Context(device=0, max_n_log2=min(24, n_log2 + 3), num_slots=1) as ctx:
        d = torch.from_numpy(np.ascontiguousarray(trace).view(np.int64)).cuda()
        for name, o in sets:
            opts = xs.ProofOptions(*o)
            try:
                proof, _ = ctx.prove_device(d.data_ptr(), n_log2, air, opts, want_times=True)
            except xs.XfgError as e:
                rows.append({"name": name, "options": list(o), "error": str(e)}); continue
            ms, launches = [], 0
            for _ in range(args.reps):
                p2, t = ctx.prove_device(d.data_ptr(), n_log2, air, opts, want_times=True)
                assert p2 == proof
                ms.append(t["device_ms"]); launches = t["kernel_launches"]
            row = {"name": name, "options": list(o), "device_ms": round(statistics.median(ms), 4), "launches": launches, "proof_bytes": len(proof)}
            # batch verification of 512 copies (tuned options: one thread block per proof, verify.cu; anything else: one thread per proof, general_verify.cuh)
            res, vt = ctx.verify_batch([proof] * 512, [air] * 512, opts, want_times=True)
            assert res == [""] * 512
            row["verify_512"] = {"kernel_ms": round(vt["kernel_ms"], 3), "total_ms": round(vt["total_ms"], 3), "proofs_per_s": round(512 / (vt["total_ms"] / 1e3))}
            if args.check:
                tr, pi, ac = orc.synthetic_case(1 << n_log2, 0)
                orc.set_threads(orc.max_threads())
                row["equals_oracle"] = proof == orc.prove(tr, pi, ac, o)
            rows.append(row)
    print(json.dumps({"tool": "options_bench", "n_log2": n_log2, "reps": args.reps, "rows": rows}))


if __name__ == "__main__":
    main()
