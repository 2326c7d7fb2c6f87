#!/usr/bin/env python
"""Regenerates profiles/ncu_traffic.json from an `ncu --set full` capture of one proof (tools/profile_round.sh):
    python tools/ncu_traffic.py gpurun_out/r02_prof.ncu-rep|r02_ncu_raw.csv [git-head]
Per kernel family of bench.py: DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, summed over the family's launches), and the ALU / FMA
pipe utilisation (time-weighted).  Families are assigned by kernel name and launch order within the proof (the capture filters the heavy kernels)."""
import csv, json, subprocess, sys, os

rep = sys.argv[1]; head = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = open(rep, errors="ignore").read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]
col = lambda n: hdr.index(n)
def num(r, n):
    try: return float(r[col(n)].replace(",", ""))
    except Exception: return 0.0
fam_seq = {"ntt1": ["ntt.interpolate_trace", "ntt.interpolate_trace", "ntt.interpolate_comp", "ntt.interpolate_comp"],
           "ntt0": ["ntt.lde_trace", "ntt.lde_trace", "ntt.lde_comp", "ntt.lde_comp"]}
seen = {"ntt1": 0, "ntt0": 0, "commit": 0}
agg = {}
for r in rows[2:]:
    name = r[col("Kernel Name")]
    if "ntt_pass_r16<1" in name or "ntt_pass_r16<(bool)1" in name or "ntt_pass_r16" in name:
        k = "ntt1" if ("ntt_pass_r16<1" in name or "ntt_pass_r16<(bool)1" in name) else "ntt0"
        seen[k] += 1
        if seen[k] > 4: continue          # launches of the next proof that the capture window still caught
        fam = fam_seq[k][seen[k] - 1]
    elif "commit_rows_kernel" in name: fam = "commit_rows.trace" if seen["commit"] == 0 else "commit_rows.comp"; seen["commit"] += 1
    elif "constraint_kernel" in name: fam = "constraints"
    elif "deep_kernel" in name: fam = "deep"
    elif "ood_kernel" in name: fam = "ood"
    elif "fri_fold" in name: fam = "fri.fold"
    elif "fri_tail" in name: fam = "fri.tail"
    elif "tree_" in name: fam = "tree"
    else: continue
    t = num(r, "gpu__time_duration.sum")
    a = agg.setdefault(fam, dict(bytes=0.0, t=0.0, alu=0.0, fma=0.0, fmah=0.0, issue=0.0, n=0))
    a["bytes"] += num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"); a["t"] += t; a["n"] += 1
    a["alu"] += t * num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") / 100
    a["fma"] += t * num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") / 100
    a["fmah"] += t * (num(r, "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active") if "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active" in hdr else 0) / 100
    a["issue"] += t * num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100
unit = rows[1][col("dram__bytes_read.sum")]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
res = {"_comment": f"per kernel family at 2^20 rows / quadratic extension, from the ncu --set full capture of one proof of build {head} (tools/profile_round.sh, tools/ncu_traffic.py): "
                   "traffic = dram__bytes_read.sum + dram__bytes_write.sum over the family's launches; pipe values = pct_of_peak_sustained_active / 100, time-weighted",
       "_git_head": head, "_launches": {k: v["n"] for k, v in agg.items()}}
for k, v in agg.items():
    res[k] = round(v["bytes"] * scale)
for key, f in (("_alu_pipe_busy", "alu"), ("_fma_pipe_busy", "fma"), ("_fmaheavy_pipe_busy", "fmah"), ("_issue_active", "issue")):
    vals = {k: round(v[f] / v["t"], 3) for k, v in agg.items() if v["t"] > 0}
    if any(vals.values()):
        res[key] = vals
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
