"""CPU tests (-m "not gpu") of the N > 1 plumbing with the gloo backend, world size 2: proofs are sharded by index with no
data-path collective; the only collectives are the barrier and the max-over-ranks of the timed region."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import xfg_stark_b200 as xs, orc
from xfg_stark_b200 import multi
rank, world, _ = multi.rank_world()
multi.init("gloo")
mine = multi.proof_indices_for_rank(5, rank, world)
# each rank "proves" its own shard (CPU oracle here; the CUDA backend on the GPU box) and only digests travel
import hashlib
digs = {}
for i in mine:
    tr, pi, ac = orc.synthetic_case(64, i)
    air = xs.pack_inputs(**{k: v for k, v in zip(("burn_amount","mint_amount","tx_prefix_hash","recipient_address","secret","network_id","target_chain_id","commitment_version"), xs.synthetic_inputs(i).values())})
    assert list(air.pub_inputs) == [int(v) for v in pi]
    digs[i] = hashlib.sha256(orc.prove(tr, pi, ac)).hexdigest()
multi.barrier()
mx = multi.max_over_ranks(10.0 + rank)
tot = multi.sum_over_ranks(len(mine))
open(os.path.join(os.environ["XFG_TEST_OUT"], "rank%%d.json" %% rank), "w").write(json.dumps({"rank": rank, "mine": mine, "max": mx, "total": tot, "digs": digs}))
multi.finalize()
'''


WIDE_WORKER = r'''
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch, torch.distributed as dist
import orc
from xfg_stark_b200 import multi
rank, world, _ = multi.rank_world()
multi.init("gloo")
W, n_log2 = 8, 8
n = 1 << n_log2; N = 8 * n; wl = W // world; n_local = n // world
rng = np.random.default_rng(99)                       # every rank builds the same wide trace and keeps only its columns
trace = rng.integers(0, 1 << 63, size=(W, n), dtype=np.uint64) %% np.uint64(orc.P)
mine = trace[rank * wl:(rank + 1) * wl]
# column-sharded interpolation + LDE (oracle stands in for the CUDA kernels), coset-major [col][k][m]
lde = np.stack([orc.lde(orc.ntt(c, 1, 1)).reshape(n, 8).T for c in mine])           # (wl, 8, n)
# all-to-all: destination r gets m in [r*n_local, (r+1)*n_local) of my columns
send = np.ascontiguousarray(np.stack([lde[:, :, r * n_local:(r + 1) * n_local] for r in range(world)]))   # (world, wl, 8, n_local)
recv = np.empty_like(send)
dist.all_to_all_single(torch.from_numpy(recv.view(np.int64)), torch.from_numpy(send.view(np.int64)))
rows = recv.reshape(world * wl, 8, n_local)          # [global column][k][m_local]: source rank s contributed columns s*wl..
# rows i = 8 m + k of my range, all W columns -> leaves -> subtree root
mat = rows.transpose(0, 2, 1).reshape(W, 8 * n_local)                                  # natural row order within my range
leaves = orc.hash_rows(mat)
root, _ = orc.merkle(leaves)
r = torch.frombuffer(bytearray(root), dtype=torch.uint8)
roots = [torch.empty_like(r) for _ in range(world)]
dist.all_gather(roots, r)
final, _ = orc.merkle(torch.stack(roots).numpy())
# single-process answer
full = np.stack([orc.lde(orc.ntt(c, 1, 1)) for c in trace])
exp, _ = orc.merkle(orc.hash_rows(full))
open(os.path.join(os.environ["XFG_TEST_OUT"], "rank%%d.json" %% rank), "w").write(json.dumps({"rank": rank, "ok": final == exp}))
multi.finalize()
'''


def test_wide_trace_sharding_logic_world_size_2(tmp_path):
    """config 5 dataflow (column-sharded LDE -> all-to-all by row ranges -> row hashing -> all-gather of subtree roots) on gloo,
    with the oracle standing in for the kernels: the sharded commitment equals the single-process one."""
    script = tmp_path / "wide.py"
    script.write_text(WIDE_WORKER % (ROOT, ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script)]
    recs = run_ranks(cmd, tmp_path)
    assert len(recs) == 2 and all(r["ok"] for r in recs)


def run_ranks(cmd, tmp_path):
    """runs the torchrun command; every rank writes its record to its own file (two processes sharing one stdout pipe can interleave)"""
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "XFG_TEST_OUT": str(tmp_path)})
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads((tmp_path / f).read_text()) for f in sorted(os.listdir(tmp_path)) if f.startswith("rank")]


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_world_size_2_sharding_and_reductions(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script)]
    recs = run_ranks(cmd, tmp_path)
    assert len(recs) == 2
    by = {r["rank"]: r for r in recs}
    assert by[0]["mine"] == [0, 2, 4] and by[1]["mine"] == [1, 3]
    assert by[0]["max"] == by[1]["max"] == 11.0 and by[0]["total"] == by[1]["total"] == 5.0
    digs = {**by[0]["digs"], **by[1]["digs"]}
    assert len(set(digs.values())) == 5            # five different proofs, each proven exactly once


def test_reference_arm_prints_one_line_under_torchrun():
    """bench.py --impl reference: rank 0 alone works and prints; other ranks exit 0 silently"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--n-log2", "16", "--ext", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
