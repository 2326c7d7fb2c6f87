"""Per-kernel time of the timed region from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file ...).
python tools/launch_table.py gpurun_out/v4_launches.csv [out_subset.csv]
A proof ends with its gather_kernel launch; with `bench.py --steps 2 --warmup 3 --no-preload --headline-only` proofs 1-3 are the warm-up, 4 the e2e
warm-up (split upload), 5-6 the timed HBM-resident region (33 launches each at 2^20 rows, replayed from the whole-proof CUDA graph)."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 14 and r[0].isdigit()]
names = [re.sub(r"\(.*", "", r[4]).replace("void ", "") for r in rows]
us = [float(r[14]) / 1e3 for r in rows]
ends = [i + 1 for i, n in enumerate(names) if "gather_kernel" in n]
proofs = [(a, b) for a, b in zip([0] + ends[:-1], ends)]
print("proofs (launch counts):", [b - a for a, b in proofs])
sel = proofs[4:6]
agg = collections.OrderedDict()
for a, b in sel:
    for i in range(a, b):
        if names[i].startswith(("int_peak", "pipe_probe")):
            break
        e = agg.setdefault(names[i], [0, 0.0]); e[0] += 1; e[1] += us[i]
tot = sum(v[1] for v in agg.values()) / len(sel)
print("| kernel | launches/proof | us/proof | share |\n|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {n} | {c // len(sel)} | {t / len(sel):.1f} | {100 * t / len(sel) / tot:.1f} % |")
print(f"| total | {sum(v[0] for v in agg.values()) // len(sel)} | {tot:.1f} | |")
if len(sys.argv) > 2:
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f); w.writerow(["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum [ns]"])
        for a, b in sel:
            for i in range(a, b):
                w.writerow([rows[i][0], rows[i][4], rows[i][7], rows[i][8], rows[i][14]])
