#!/usr/bin/env python
"""bench.py — the burn-mint proving benchmark (BASELINE.json: "burn-mint proof latency (ms) @2^20 trace").

One "step" = one complete BurnMintAir proof (trace -> StarkProof bytes) of the named workload on one GPU:
normalised BurnMintAir, 2^20 rows x 7 columns, blowup 8, quadratic extension, 42 queries, grinding 4, FRI folding 8,
remainder max degree 31 (BASELINE config 3; synthetic inputs of SURVEY.md §8d).

  python bench.py [--gpus N] [--steps K] [--warmup W]           # this repo's CUDA backend through the C ABI
  python bench.py --impl reference [...]                        # the CPU oracle (restatement of the reference's Winterfell
                                                                # path; the Rust crate cannot be built here) on the host cores

Under torchrun (N > 1) every rank drives its own GPU with its own independent proofs (replicas, no collective on the data
path: SURVEY.md §8e); the timed region is bracketed by a barrier + synchronize and the max over ranks is reported.
`value` = milliseconds per proof over the whole job (time of the K-step region / proofs completed by all ranks) with the
trace already resident in HBM; `e2e` = the same through xfg_prove_burn_mint with the trace in pinned HOST memory (upload
and proof download inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "burn-mint proof latency (ms) @2^20 trace"
W_COLS, BLOWUP, CE = 7, 8, 2


def algorithmic_bytes(n_log2, e):
    """each input read once, each output written once; per kernel family of this backend."""
    n = 1 << n_log2; N = BLOWUP * n; w = W_COLS; c = CE
    fri = 0; nl = N
    layers = 0
    while nl > 256:                       # (rem_max_deg + 1) * blowup = 256
        fri += 8 * e * nl + (32 + 96 + 8 * e) * nl // 8; nl //= 8; layers += 1
    return {
        "ntt.interpolate_trace": 16 * w * n,
        "ntt.lde_trace": 8 * w * n + 8 * w * N,
        "commit_rows.trace": 8 * w * N + 32 * N + 96 * (N - N // 8),       # leaves + the 3 fused tree levels
        "tree_upper.trace": 96 * (N // 8),
        "constraints": 8 * w * c * n + 8 * e * c * n,
        "ntt.interpolate_comp": 16 * e * c * n,
        "combine": 8 * e * c * n + 8 * e * n,
        "ntt.lde_comp": 8 * e * (n + N),
        "commit_rows.comp": 8 * e * N + 32 * N + 96 * (N - N // 8),
        "tree_upper.comp": 96 * (N // 8),
        "ood": 8 * w * n + 8 * e * n,
        "deep": 8 * w * N + 8 * e * N + 8 * e * N + 32 * (N // 8),
        "fri.fold": sum(8 * e * (N >> (3 * l)) + 8 * e * (N >> (3 * l + 3)) for l in range(layers)),
        "fri.tree": sum(96 * (N >> (3 * l + 3)) for l in range(layers)),
        "_total_survey": 16 * w * n + 8 * w * n + 8 * w * N + 8 * w * N + 32 * N + 96 * N + 8 * w * c * n + 8 * e * c * n + 16 * e * c * n
                         + 8 * e * (n + N) + 8 * e * N + 32 * N + 96 * N + 8 * w * n + 8 * e * n + 8 * w * N + 8 * e * N + 8 * e * N + fri,
    }


def ntt_gbps_from(acc, n_log2, ext):
    """BASELINE.json's third metric, "NTT GB/s" = 16 B x elements / time of one batched transform, per NTT family of the proof.
    acc: {family: [ms, launches]} from the per-kernel profile."""
    n = 1 << n_log2
    out = {}
    for name, transforms in (("ntt.interpolate_trace", W_COLS), ("ntt.lde_trace", W_COLS * BLOWUP), ("ntt.interpolate_comp", CE * ext), ("ntt.lde_comp", 6 * ext)):
        ms = acc.get(name, [0.0, 0])[0]
        if ms > 0:
            out[name] = {"transforms": transforms, "gbps": round(16.0 * transforms * n / ms / 1e6, 1)}
    return out


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True); self.t.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 3.0:      # nvidia-smi needs ~0.3 s to deliver its first sample
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25); self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    """Host threads the CPU arm uses: every core this process may run on, regardless of OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_reference_ms(n_log2, ext, budget_s, steps=1):
    """Times the CPU oracle on the FULL workload (one complete 2^n_log2-row proof per step, never a smaller trace) with every host core.
    The budget only bounds HOW MANY proofs are timed.  Returns (ms per proof list, cores, sample text)."""
    import orc
    cores = host_cores(); orc.set_threads(cores)
    opts = (42, 8, 4, ext, 8, 31)
    tr, pi, ac = orc.synthetic_case(1 << n_log2, 0)
    ts = []; t_start = time.perf_counter()
    for k in range(max(1, steps)):
        t0 = time.perf_counter(); orc.prove(tr, pi, ac, opts); ts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start + ts[-1] > budget_s:
            break
    txt = (f"{len(ts)} full 2^{n_log2}-row proof(s), one per step, ext degree {ext}, {cores} OpenMP threads (all host cores; OMP_NUM_THREADS ignored), "
           "oracle restatement of the reference's Winterfell path (the Rust crate cannot be built here)")
    return [t * 1e3 for t in ts], cores, txt


def run_reference(args, rank, world):
    if rank != 0:
        return
    for _ in range(args.warmup):
        pass   # warm-up has no meaning for the CPU arm beyond page-faulting the oracle: the first timed call builds its twiddle caches
    ms, cores, sample = cpu_reference_ms(args.n_log2, args.ext, budget_s=200.0, steps=max(1, args.steps))
    v = statistics.mean(ms)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": v, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks) + u32 ARX (BLAKE3)",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def measure_batch(args, rank, world, local):
    """BASELINE config 4: `batch_total` independent 2^16-row proofs (no extension), proof i on GPU i mod G, each GPU pipelining
    its share over `slots` streams through xfg_prove_burn_mint_batch (host traces in, proof bytes out).  Strong scaling.
    Returns the result dict (every rank)."""
    import numpy as np
    import torch
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import multi
    n_log2 = 16
    opts = xs.ProofOptions()
    ctx = xs.Context(device=local, max_n_log2=n_log2, num_slots=args.slots)
    mine = multi.proof_indices_for_rank(args.batch_total, rank, world)
    distinct = 32                                                    # 32 distinct synthetic inputs per rank, cycled (bounds host memory)
    airs, traces, keep = [], [], []
    for k in range(distinct):
        s = xs.synthetic_inputs(rank * distinct + k)
        a = ctx.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
        pinned = torch.empty((7, 1 << n_log2), dtype=torch.int64).pin_memory()        # traces live in pinned host memory (uploaded without staging)
        view = pinned.numpy().view(np.uint64); view[:] = ctx.build_trace(a, n_log2)
        airs.append(a); traces.append(view); keep.append(pinned)
    tl = [traces[i % distinct] for i in range(len(mine))]; al = [airs[i % distinct] for i in range(len(mine))]
    ctx.prove_batch(tl[:8], al[:8], opts)                            # warm-up
    best = None
    for _ in range(max(1, min(args.steps, 3))):
        multi.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        proofs, dev_ms = ctx.prove_batch(tl, al, opts)
        torch.cuda.synchronize(); multi.barrier()
        ms = multi.max_over_ranks((time.perf_counter() - t0) * 1e3, device="cuda")
        best = ms if best is None else min(best, ms)
    assert all(len(p) > 1000 for p in proofs)
    ctx.close()
    return {"metric": "burn-mint proofs/s (1024 x 2^16-row proofs)", "value": args.batch_total / (best / 1e3), "unit": "proofs/s", "n_gpus": world,
            "higher_is_better": True, "scaling": "strong", "batch_total": args.batch_total, "wall_ms": best, "slots": args.slots,
            "h2d_bytes_per_proof": 7 * 8 << n_log2, "proof_bytes": len(proofs[0]),
            "config": {"workload": "1024 independent BurnMintAir proofs, 2^16 rows, blowup 8, no extension (BASELINE config 4), host traces in / proof bytes out"}}


def run_batch(args, rank, world, local):
    from xfg_stark_b200 import multi
    out = measure_batch(args, rank, world, local)
    if rank == 0:
        print(json.dumps(out))
    multi.finalize()


def run_verify(args, rank, world, local):
    """`batch_total` proofs (2^16 rows, no extension: the proofs of BASELINE config 4) verified by
    xfg_verify_burn_mint_batch, proof i on GPU i mod G.  `value` counts the whole call (host walk of the length prefixes, staging copy,
    H2D, kernel, D2H); the kernel-only rate and the CPU oracle verifier (one host thread, a bounded sample) are reported beside it."""
    import numpy as np
    import torch
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import multi
    n_log2 = 16
    opts = xs.ProofOptions()
    ctx = xs.Context(device=local, max_n_log2=n_log2, num_slots=4)
    mine = multi.proof_indices_for_rank(args.batch_total, rank, world)
    distinct = 16
    airs, traces = [], []
    for k in range(distinct):
        s = xs.synthetic_inputs(rank * distinct + k)
        a = ctx.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
        airs.append(a); traces.append(ctx.build_trace(a, n_log2))
    base_proofs, _ = ctx.prove_batch(traces, airs, opts)
    pl = [base_proofs[i % distinct] for i in range(len(mine))]; al = [airs[i % distinct] for i in range(len(mine))]
    assert ctx.verify_batch(pl[:8], al[:8], opts) == [""] * 8          # warm-up
    best, best_t = None, None
    for _ in range(max(3, args.steps)):
        multi.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, vt = ctx.verify_batch(pl, al, opts, want_times=True)
        ms = multi.max_over_ranks((time.perf_counter() - t0) * 1e3, device="cuda")
        if best is None or ms < best:
            best, best_t = ms, vt
    assert res == [""] * len(pl)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc                                                       # CPU baseline leg only
        pis = []
        for k in range(4):
            _, pi, ac = orc.synthetic_case(1 << n_log2, k); pis.append((pi, ac))
        t0 = time.perf_counter(); cnt = 0
        while time.perf_counter() - t0 < 5.0:
            for k in range(4):
                assert orc.verify(base_proofs[k], pis[k][0], pis[k][1], opts.as_tuple()) == ""; cnt += 1
        cpu = {"value": cnt / (time.perf_counter() - t0), "unit": "proofs/s", "cores": 1, "kind": "port", "sample": f"{cnt} verifications of 2^16-row proofs by the CPU oracle verifier, one thread"}
    if rank == 0:
        print(json.dumps({"metric": "burn-mint proofs verified/s (2^16-row proofs)", "value": args.batch_total / (best / 1e3), "unit": "proofs/s", "n_gpus": world,
                          "higher_is_better": True, "scaling": "strong", "batch_total": args.batch_total, "wall_ms": best, "proof_bytes": len(base_proofs[0]),
                          "kernel_only_proofs_per_s": len(pl) / (best_t["kernel_ms"] / 1e3), "times_rank0": best_t, "gpu_launches": 1, "cpu_baseline": cpu,
                          "config": {"workload": "batch verification of independent BurnMintAir proofs, 2^16 rows, blowup 8, no extension, 42 queries; proof bytes in host memory, verdicts out"}}))
    ctx.close(); multi.finalize()


def run_air(args, rank, world, local):
    """one proof of the "wide synthetic AIR" with real constraints through the generic front-end
    (xfg_prove_air): `air_width` registers, x_j' = x_j * x_(j+1) + c_j (degree 2), 2^n rows.  Replicas across GPUs, as the latency workload."""
    import numpy as np
    import torch
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import multi, air as A
    n_log2 = args.n_log2 if args.n_log2 != 20 else 16
    W = args.air_width
    opts = xs.ProofOptions(field_extension=args.ext)
    air, trace = A.wide_quadratic_air(W, 1 << n_log2, seed=1 + rank, extra_steps=(1 << (n_log2 - 1),))
    ctx = xs.Context(device=local, max_n_log2=n_log2, num_slots=1, max_width=W)
    h_trace = torch.empty((W, 1 << n_log2), dtype=torch.int64).pin_memory()
    h_np = h_trace.numpy().view(np.uint64); h_np[:] = trace
    d_trace = h_trace.cuda(); torch.cuda.synchronize()
    dev_fn = lambda: ctx.prove_air_device(air, d_trace.data_ptr(), n_log2, opts)
    e2e_fn = lambda: ctx.prove_air(air, h_np, opts)

    def region(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        multi.barrier(); torch.cuda.synchronize(); e0.record()
        for _ in range(steps):
            last = fn()
        e1.record(); e1.synchronize(); multi.barrier()
        return multi.max_over_ranks(e0.elapsed_time(e1), device="cuda"), last
    for _ in range(args.warmup):
        dev_fn()
    e2e_fn()
    sampler = ClockSampler(local); sampler.start()
    t_end = time.time() + 0.6
    while time.time() < t_end:
        dev_fn()
    dev_ms, proof = region(dev_fn, args.steps)
    e2e_ms, proof2 = region(e2e_fn, args.steps)
    clocks = sampler.stop()
    assert proof == proof2
    _, times = ctx.prove_air_device(air, d_trace.data_ptr(), n_log2, opts, want_times=True)
    _, times2 = ctx.prove_air(air, h_np, opts, want_times=True)
    ctx.set_profiling(True); acc = {}
    for _ in range(3):
        ctx.prove_air_device(air, d_trace.data_ptr(), n_log2, opts, want_times=True)
        for name, ms, launches in ctx.get_profile():
            a = acc.setdefault(name, [0.0, 0]); a[0] += ms / 3.0; a[1] = launches
    ctx.set_profiling(False)
    n = 1 << n_log2; N = 8 * n; e = args.ext
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        peak, peak_src = 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"
    ab = {"ntt.interpolate_trace": 16 * W * n, "ntt.lde_trace": 8 * W * n + 8 * W * N, "commit_rows.trace": 8 * W * N + 32 * N + 96 * (N - N // 8),
          "constraints": 8 * W * 2 * n + 8 * e * 2 * n, "deep": 8 * W * N + 16 * e * N + 4 * N, "ood": 8 * W * n + 8 * e * n,
          "commit_rows.comp": 8 * e * N + 32 * N + 96 * (N - N // 8), "ntt.lde_comp": 8 * e * (n + N)}
    kernels = [{"name": k, "ms": round(ms, 4), "launches": l, "alg_bytes": ab.get(k), "gbps": round(ab[k] / ms / 1e6, 1) if k in ab and ms > 0 else None,
                "frac": round(ab[k] / ms / 1e6 / peak, 4) if k in ab and ms > 0 else None} for k, (ms, l) in sorted(acc.items(), key=lambda kv: -kv[1][0])]
    top = kernels[0]
    out = {"metric": f"generic-AIR proof latency (ms): {W} registers x 2^{n_log2} rows, degree-2 constraints", "value": dev_ms / (args.steps * world), "unit": "ms",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": False, "scaling": "weak",
           "vs_baseline": None, "dtype": "u64 (Goldilocks) + u32 ARX (BLAKE3)", "data": "synthetic",
           "config": {"workload": f"wide synthetic AIR through the generic front-end (xfg_prove_air): {W} registers, x_j' = x_j x_(j+1) + c_j, 2^{n_log2} rows, blowup 8, "
                                  f"{'quadratic' if e == 2 else 'no'} extension, 42 queries, {W + 2} assertions on 3 steps", "l2": "working set >> L2 at the default size; no explicit flush"},
           "proof_bytes": len(proof), "clocks": clocks,
           "e2e": {"value": e2e_ms / (args.steps * world), "unit": "ms", "h2d_bytes_per_step": times2["h2d_bytes"], "d2h_bytes_per_step": times2["d2h_bytes"]},
           "other_options": other_options,
        "gpu_launches": times["kernel_launches"] * args.steps, "device_ms_per_proof": times["device_ms"],
           "stages_ms": {k: round(v, 4) for k, v in times.items() if k in xs.STAGE_NAMES},
           "roofline": {"bound": "hbm", "kernel": top["name"], "achieved": top["gbps"], "peak": peak, "unit": "GB/s", "frac": top["frac"], "traffic": None,
                        "peak_source": peak_src, "launch_ms": top["ms"], "alg_bytes": top["alg_bytes"],
                        "note": "integer-pipe bound like the burn-mint path (DESIGN.md section 4)"},
           "kernels": kernels}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import orc                                                       # CPU baseline leg only
        cores = orc.max_threads(); orc.set_threads(cores)
        f = air.flatten(); t0 = time.perf_counter(); ref = orc.prove_air(f, trace, opts.as_tuple()); cpu_ms = (time.perf_counter() - t0) * 1e3
        orc.set_threads(1)
        assert ref == proof, "GPU proof differs from the CPU oracle"
        out["cpu_baseline"] = {"value": cpu_ms, "unit": "ms", "cores": cores, "kind": "port", "sample": f"the same proof once by the CPU oracle's generic prover, {cores} OpenMP threads (bytes compared with the GPU proof)"}
    if rank == 0:
        print(json.dumps(out))
    ctx.close(); multi.finalize()


def workload_config(args):
    return {"workload": f"BurnMintAir synthetic trace 2^{args.n_log2} rows x 7 cols, blowup 8, {'quadratic' if args.ext == 2 else 'no'} extension, "
                        f"42 queries, grinding 4, FRI folding 8, remainder max degree 31 (BASELINE config {'3' if args.n_log2 == 20 else '2-like'})",
            "parallelism": "one independent proof stream per GPU (replicas, no collective)",
            "l2": "working set ~2 KB per trace row (2.1 GB at 2^20) >> 126 MB L2; no explicit flush between steps",
            "aggregate": "value = ms per proof over the whole job = time of the K-step region / (K * n_gpus)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-log2", type=int, default=20)
    ap.add_argument("--ext", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-preload", action="store_true", help="skip the 0.6 s untimed load loop before the timed region (deterministic launch count for ncu)")
    ap.add_argument("--workload", default="latency", choices=["latency", "batch", "wide", "verify", "air"],
                    help="latency: the headline (one 2^n proof per step); batch: BASELINE config 4, 1024 independent 2^16 proofs sharded over "
                         "the GPUs; wide: BASELINE config 5, one 64-column x 2^24-row trace, column-sharded LDE with the all-to-all fused into the "
                         "last NTT pass, then row hashing (needs --gpus >= 2 for a real exchange)")
    ap.add_argument("--headline-only", action="store_true", help="latency workload: skip the batch (config 4) and wide (config 5, N >= 2) measurements that ride on the same JSON line")
    ap.add_argument("--wide-log2", type=int, default=24); ap.add_argument("--wide-check-log2", type=int, default=20)
    ap.add_argument("--air-width", type=int, default=64, help="registers of the wide synthetic AIR for --workload air (generic front-end; 2^16 rows unless --n-log2 is given)")
    ap.add_argument("--batch-total", type=int, default=1024)
    ap.add_argument("--slots", type=int, default=16, help="proof workspaces/streams per GPU for --workload batch (1487 / 2148 / 2784 / 3145 proofs/s at 2 / 4 / 8 / 16 on one B200)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "wide":
        import wide_worker
        wide_worker.run(args.n_log2 if args.n_log2 != 20 else 24, 64, steps=max(1, args.steps))
        return

    import numpy as np
    import torch
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import multi

    # keep stdout to the one JSON line: NCCL writes its banner / debug output to stdout unless told otherwise
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL_DEBUG itself is left as the caller set it (the driver reads the rank count from its log)
    torch.cuda.set_device(local)
    multi.init("nccl", torch.device("cuda", local))
    if args.workload == "batch":
        return run_batch(args, rank, world, local)
    if args.workload == "verify":
        return run_verify(args, rank, world, local)
    if args.workload == "air":
        return run_air(args, rank, world, local)
    opts = xs.ProofOptions(field_extension=args.ext)
    ctx = xs.Context(device=local, max_n_log2=args.n_log2, num_slots=1)
    n = 1 << args.n_log2
    s = xs.synthetic_inputs(rank)
    air = ctx.pack_inputs(s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
    h_trace = torch.empty((7, n), dtype=torch.int64).pin_memory()           # pinned host trace (e2e input)
    h_np = h_trace.numpy().view(np.uint64)
    h_np[:] = ctx.build_trace(air, args.n_log2)
    d_trace = h_trace.cuda()                                                 # HBM-resident trace (`value` input)
    torch.cuda.synchronize()

    def barrier():
        multi.barrier()
        torch.cuda.synchronize()

    def timed_region(fn, steps):
        """K steps bracketed by barrier + synchronize; CUDA events on torch's stream (each step is host-synchronous);
        returns the max over ranks of the region time in ms, and the last proof."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(); e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record(); e1.synchronize(); barrier()
        return multi.max_over_ranks(e0.elapsed_time(e1), device="cuda"), last

    # timed calls carry no per-stage events (the library then replays its whole-proof CUDA graph where it can); the stage and
    # per-kernel breakdowns come from separate instrumented calls after the timed regions
    dev_fn = lambda: ctx.prove_device(d_trace.data_ptr(), args.n_log2, air, opts)
    e2e_fn = lambda: ctx.prove(h_np, air, opts)
    dev_fn_timed = lambda: ctx.prove_device(d_trace.data_ptr(), args.n_log2, air, opts, want_times=True)
    for _ in range(args.warmup):
        dev_fn()
    e2e_fn()
    sampler = ClockSampler(local); sampler.start()
    t_end = time.time() + (0.0 if args.no_preload else 0.6)
    while time.time() < t_end:            # keep the GPU under the same load until the sampler has a few readings (untimed)
        dev_fn()
    dev_ms, proof = timed_region(dev_fn, args.steps)
    e2e_ms, proof2 = timed_region(e2e_fn, args.steps)
    # the reference's real callers (SURVEY.md 8b / src/burn_mint_prover.rs:62-129), same timed-region rules:
    #  from_inputs : the 8-argument entry, trace built on the device (0 B of trace upload)
    #  mont_cols   : seven registered column buffers in Montgomery form (TraceTable::get_column memory), no conversion pass
    #  pageable    : the contiguous canonical trace in ordinary (pageable) host memory, staged by the library
    inp = (s["burn"], s["mint"], s["tx_prefix_hash"], s["recipient"], s["secret"], s["network_id"], s["target_chain_id"], s["version"])
    callers = {}
    try:        # auxiliary measurements: a failure here is reported on the line, it must not cost the headline numbers
        fi_fn = lambda: ctx.prove_from_inputs(*inp, n_log2=args.n_log2, options=opts)
        pg_np = np.array(h_np, copy=True)
        pg_fn = lambda: ctx.prove(pg_np, air, opts)
        R = (1 << 64) % xs.P
        mont = [np.full(n, (int(h_np[c, 0]) * R) % xs.P, dtype=np.uint64) if c != 4 else
                np.array([(v * R) % xs.P for v in range(4)], dtype=np.uint64)[h_np[4].astype(np.int64)] for c in range(7)]
        for arr in mont:
            ctx.host_register(arr)
        mc_fn = lambda: ctx.prove_cols(mont, air, opts, form=1)
        for name, fn in (("from_inputs", fi_fn), ("mont_cols", mc_fn), ("pageable", pg_fn)):
            fn()
            ms_c, pr = timed_region(fn, args.steps)
            assert pr == proof, f"{name}: proof differs"
            callers[name] = ms_c / (args.steps * world)
        for arr in mont:
            ctx.host_unregister(arr)
    except AssertionError:
        raise
    except Exception as e:
        callers["error"] = repr(e)
    # `with_options` (src/burn_mint_prover.rs:44-49) outside the reference's own setting: the general-options pipeline on the same device-resident
    # trace (auxiliary: folding factor 4, and the cubic extension); their bytes are checked against the oracle by tests/test_gpu_options.py
    other_options = {}
    try:
        for name, o in (("blowup8_folding4_quadratic", (42, 8, 4, 2, 4, 31)), ("blowup8_folding8_cubic", (42, 8, 4, 3, 8, 31))):
            oo = xs.ProofOptions(*o)
            oo_fn = lambda: ctx.prove_device(d_trace.data_ptr(), args.n_log2, air, oo)
            oo_fn()
            k_o = max(3, args.steps // 3)
            ms_o, pr_o = timed_region(oo_fn, k_o)
            other_options[name] = {"options": list(o), "ms_per_proof": ms_o / (k_o * world), "steps": k_o, "proof_bytes": len(pr_o)}
    except Exception as e:
        other_options["error"] = repr(e)
    clocks = sampler.stop()
    assert proof == proof2, "device-resident and host-buffer proofs differ"
    # sustained: >= 2.5 s of back-to-back proofs with its own clock samples (thermal / power behaviour of a long batch of large proofs)
    sus_steps = max(args.steps, int(2500.0 / max(dev_ms / args.steps, 1e-3)))
    sampler2 = ClockSampler(local); sampler2.start()
    sus_ms, proof_s = timed_region(dev_fn, sus_steps)
    sus_clocks = sampler2.stop()
    assert proof_s == proof
    proof3, times = dev_fn_timed()
    _, times2 = ctx.prove(h_np, air, opts, want_times=True)
    assert proof3 == proof

    # per-kernel-family device times (CUDA events on the library's stream around each launcher), 3 profiled proofs
    ctx.set_profiling(True)
    acc = {}
    for _ in range(3):
        _, t = dev_fn_timed()
        for name, ms, launches in ctx.get_profile():
            a = acc.setdefault(name, [0.0, 0]); a[0] += ms / 3.0; a[1] = launches
    ctx.set_profiling(False)
    ab = algorithmic_bytes(args.n_log2, args.ext)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kernels = []
    for name, (ms, launches) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        b = ab.get(name)
        kernels.append({"name": name, "ms": round(ms, 4), "launches": launches, "alg_bytes": b,
                        "gbps": round(b / ms / 1e6, 1) if b and ms > 0 else None, "frac": round(b / ms / 1e6 / peak, 4) if b and ms > 0 else None})
    top = kernels[0]
    traffic, alu_busy = None, None
    try:
        nt = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = nt.get(top["name"]); alu_busy = nt.get("_alu_pipe_busy", {}).get(top["name"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top["name"], "achieved": top["gbps"], "peak": peak, "unit": "GB/s", "frac": top["frac"], "traffic": traffic,
                "peak_source": peak_src, "launch_ms": top["ms"], "alg_bytes": top["alg_bytes"],
                "alu_pipe_busy_ncu": alu_busy,
                "note": "every heavy kernel of this path is bound by the 32-bit integer pipes (64-bit modular arithmetic, BLAKE3), not by HBM: ncu ALU pipe 56-93 % busy, DRAM 8-25 % (DESIGN.md section 4)",
                "whole_proof": {"alg_bytes": ab["_total_survey"], "gbps": round(ab["_total_survey"] / times["device_ms"] / 1e6, 1),
                                "frac": round(ab["_total_survey"] / times["device_ms"] / 1e6 / peak, 4)}}

    # integer-pipe roofline of the BLAKE3 kernels: measured ALU-pipe peak (LOP3/SHF microbenchmark, no memory
    # traffic) against compressions/s x 458 ALU-pipe instructions per compression (SASS of the compression loop of commit_rows_kernel: 8 per G x 56 G + 10;
    # the 4 additions of G run on the FMA pipe as IMAD)
    int_peak = ctx.int_pipe_peak()
    n_rows = 1 << args.n_log2; ALU_PER_COMPRESSION = 458
    fri_leaves = []; nl = 8 * n_rows
    while nl > 256:
        fri_leaves.append(nl // 8); nl //= 8
    compressions = {"commit_rows.trace": 15 * n_rows, "commit_rows.comp": 15 * n_rows, "tree_upper.trace": n_rows - 1, "tree_upper.comp": n_rows - 1,
                    "fri.tree": sum(r - 1 for r in fri_leaves)}
    int_pipe = {"peak_gops": round(int_peak, 1), "unit": "1e9 ALU-pipe instructions/s (measured, LOP3+SHF chains)", "alu_instr_per_compression": ALU_PER_COMPRESSION, "kernels": {}}
    for name, (ms, _) in acc.items():
        if name in compressions and ms > 0:
            g = compressions[name] * ALU_PER_COMPRESSION / ms / 1e6
            int_pipe["kernels"][name] = {"compressions": compressions[name], "gcompressions_per_s": round(compressions[name] / ms / 1e6, 2),
                                         "achieved_gops": round(g, 1), "frac": round(g / int_peak, 4)}

    out = {
        "metric": METRIC, "value": dev_ms / (args.steps * world), "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 (Goldilocks) + u32 ARX (BLAKE3)", "data": "synthetic", "config": workload_config(args),
        "proofs_per_s": args.steps * world / (dev_ms / 1e3), "proof_bytes": len(proof),
        "clocks": clocks,
        "sustained": {"ms_per_proof": sus_ms / (sus_steps * world), "ms_per_step": sus_ms / sus_steps, "steps": sus_steps, "seconds": sus_ms / 1e3, "clocks": sus_clocks},
        "e2e": {"value": e2e_ms / (args.steps * world), "unit": "ms", "h2d_bytes_per_step": times2["h2d_bytes"], "d2h_bytes_per_step": times2["d2h_bytes"],
                "proofs_per_s": args.steps * world / (e2e_ms / 1e3)},
        "e2e_callers_ms": {"pinned_contiguous_canonical": e2e_ms / (args.steps * world), "from_inputs_device_built_trace": callers.get("from_inputs"),
                           "registered_montgomery_columns": callers.get("mont_cols"), "pageable_contiguous_canonical": callers.get("pageable"), "error": callers.get("error"),
                           "note": "same proof bytes on every path (asserted); from_inputs uploads only the 1 KB init block of the proof state, the others also the 7 x n x 8 B trace"},
        "other_options": other_options,
        "gpu_launches": times["kernel_launches"] * args.steps,
        "device_ms_per_proof": times["device_ms"],
        "stages_ms": {k: round(v, 4) for k, v in times.items() if k in xs.STAGE_NAMES},
        "e2e_stages_ms": dict({k: round(v, 4) for k, v in times2.items() if k in xs.STAGE_NAMES}, h2d_until_first_kernel_ms=round(times2["h2d_ms"], 4),
                              device_ms=round(times2["device_ms"], 4), total_ms=round(times2["total_ms"], 4)),
        "roofline": roofline, "int_pipe": int_pipe, "kernels": kernels,
    }
    try:
        out["ntt_gbps"] = ntt_gbps_from(acc, args.n_log2, args.ext)
    except Exception as e:          # reporting only: never lose the bench line over it
        out["ntt_gbps"] = {"error": str(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ms, cores, sample = cpu_reference_ms(args.n_log2, args.ext, budget_s=30.0, steps=1)
        out["cpu_baseline"] = {"value": ms[0], "unit": "ms", "cores": cores, "kind": "port", "sample": sample}
        # the reference's actual build is single-threaded: one thread on a 2^16-row proof, scaled n log n
        import orc
        orc.set_threads(1)
        tr, pi, ac = orc.synthetic_case(1 << 16, 0)
        t0 = time.perf_counter(); orc.prove(tr, pi, ac, (42, 8, 4, args.ext, 8, 31)); t1 = (time.perf_counter() - t0) * 1e3
        scale = (1 << (args.n_log2 - 16)) * args.n_log2 / 16.0 if args.n_log2 > 16 else 1.0
        out["cpu_baseline"]["single_thread"] = {"value": t1 * scale, "unit": "ms", "cores": 1,
                                                "sample": (f"2^16-row proof on one thread ({t1:.0f} ms)" + (f", scaled x{scale:.1f} (n log n) to 2^{args.n_log2}" if scale != 1.0 else ""))}
    ctx.close()
    # BASELINE's other two configurations ride on the same line so that the driver's per-N records carry them: config 4 (1024 x 2^16 batch,
    # every N) and, where a real exchange exists (N >= 2), config 5 (64 x 2^24 wide trace; checked against the oracle at 64 x 2^20 first)
    if not args.headline_only:
        try:
            out["batch"] = measure_batch(args, rank, world, local)
        except Exception as e:
            out["batch"] = {"error": repr(e)}
        if world >= 2:
            import wide_worker
            try:
                chk = wide_worker.measure(args.wide_check_log2, 64, steps=1, check=True)
                w = wide_worker.measure(args.wide_log2, 64, steps=2)
                w["check"] = chk["check"]; w["check_workload"] = chk["config"]["workload"]; w["check_root"] = chk["root"]; w["check_ms"] = chk["value"]
                out["wide"] = w
            except Exception as e:
                out["wide"] = {"error": repr(e), "check": False}
    if rank == 0:
        print(json.dumps(out))
    multi.finalize()


if __name__ == "__main__":
    main()
