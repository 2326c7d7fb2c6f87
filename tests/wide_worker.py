"""One process per GPU (torchrun): BASELINE config 5 through the C ABI.  Used by tests (--check: compare with the oracle) and
by bench.py --workload wide.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def make_columns(n, W, world, rank, seed, torch):
    """Synthetic wide trace: uniformly random canonical elements (values < 2^63 < p) generated on the device.  Seeded per GLOBAL
    column, so the trace - and therefore the commitment - is the same for every number of ranks."""
    wl = W // world
    cols = torch.empty((wl, n), dtype=torch.int64, device="cuda")
    g = torch.Generator(device="cuda")
    for c in range(wl):
        g.manual_seed(seed + rank * wl + c)
        cols[c] = torch.randint(0, (1 << 63) - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    return cols


def measure(n_log2, W, steps=1, check=False, seed=1234):
    """Config 5 on the already initialised process group (one process per GPU); returns the result dict on every rank."""
    import torch
    import torch.distributed as dist
    import xfg_stark_b200 as xs
    from xfg_stark_b200 import multi
    rank, world, local = multi.rank_world()
    n = 1 << n_log2; wl = W // world
    ctx = xs.Context(device=local, max_n_log2=min(n_log2, 16), num_slots=1)
    wide = xs.WideTrace(ctx, n_log2, W, world, rank)
    cols = make_columns(n, W, world, rank, seed, torch)
    # exchange IPC handles of the receive buffers (bytes over NCCL), then map the peers
    h = torch.frombuffer(bytearray(wide.ipc_handle()), dtype=torch.uint8).cuda()
    allh = [torch.empty_like(h) for _ in range(world)]
    if world > 1:
        dist.all_gather(allh, h)
        wide.open_peers([bytes(t.cpu().numpy().tobytes()) for t in allh])
    best = None
    for _ in range(max(1, steps)):
        multi.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ext_ms = wide.extend(cols.data_ptr())      # interpolate + LDE + fused all-to-all (returns after this rank's stores are issued and complete)
        multi.barrier()                            # every rank's peer stores have landed
        root, commit_ms = wide.commit()
        r = torch.frombuffer(bytearray(root), dtype=torch.uint8).cuda()
        roots = [torch.empty_like(r) for _ in range(world)]
        if world > 1:
            dist.all_gather(roots, r)
            final = ctx.merkle_root(torch.stack(roots).cpu().numpy())
        else:
            final = root
        e1.record(); e1.synchronize()
        ms = multi.max_over_ranks(e0.elapsed_time(e1), device="cuda")
        ext = multi.max_over_ranks(ext_ms, device="cuda"); com = multi.max_over_ranks(commit_ms, device="cuda")
        if best is None or ms < best[0]:
            best = (ms, ext, com)
    ok = None
    if check:
        full = [torch.empty_like(cols) for _ in range(world)]
        if world > 1:
            dist.all_gather(full, cols)
        else:
            full = [cols]
        if rank == 0:
            import orc
            orc.set_threads(max(1, len(os.sched_getaffinity(0))))
            trace = torch.cat(full).cpu().numpy().view(np.uint64)
            lde = np.stack([orc.lde(orc.ntt(trace[c], 1, 1)) for c in range(W)])
            exp_root, _ = orc.merkle(orc.hash_rows(lde))
            ok = bool(exp_root == final)
        flag = torch.tensor([1 if (ok or rank != 0) else 0], device="cuda")
        if world > 1:
            dist.broadcast(flag, 0)
        ok = bool(flag.item())
    N = 8 * n
    sent = (world - 1) / world * W * N * 8 / world                      # bytes each rank stores into OTHER ranks' buffers
    out = {"metric": "wide-trace LDE + row commitment (ms)", "value": best[0], "unit": "ms", "n_gpus": world, "higher_is_better": False,
           "extend_ms": best[1], "commit_ms": best[2], "ntt_gbps_per_gpu": 16.0 * wl * n * 9 / best[1] / 1e6,
           "peer_store_gbps_per_gpu": sent / best[1] / 1e6, "root": final.hex(), "check": ok, "seed": seed,
           "config": {"workload": f"one {W}-column x 2^{n_log2}-row random trace (seeded per global column: same root for any number of GPUs), blowup 8, "
                                  f"columns sharded over {world} GPUs, all-to-all fused into the last NTT pass (peer stores), BLAKE3 row hashing + Merkle (BASELINE config 5)"}}
    wide.close(); ctx.close()
    del cols; torch.cuda.empty_cache()
    return out


def run(n_log2, W, steps=1, check=False, seed=1234):
    import torch
    from xfg_stark_b200 import multi
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")          # NCCL's log must not pollute the JSON line on stdout; NCCL_DEBUG itself is left alone
    rank, world, local = multi.rank_world()
    torch.cuda.set_device(local)
    multi.init("nccl", torch.device("cuda", local))
    out = measure(n_log2, W, steps, check, seed)
    if check:
        assert out["check"], "sharded commitment differs from the oracle"
    if rank == 0:
        print(json.dumps(out))
        print("WIDE_OK")
    multi.finalize()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-log2", type=int, default=24); ap.add_argument("--cols", type=int, default=64)
    ap.add_argument("--steps", type=int, default=1); ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    run(a.n_log2, a.cols, a.steps, a.check)
