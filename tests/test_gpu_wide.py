"""GPU parity tests for BASELINE config 5 (one wide trace, columns sharded over G ranks, all-to-all fused into the last NTT
pass as peer stores, row hashing + subtree per rank).  On a single GPU the G ranks are emulated in one process (their
receive buffers are plain device pointers of the same GPU), which exercises exactly the same kernels and index arithmetic;
`test_wide_two_processes_ipc` runs the real one-process-per-GPU path over CUDA IPC when two GPUs are present."""
import os
import subprocess
import sys

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_wide(trace):
    """-> (LDE (W, 8n) natural order, leaf digests (8n, 32), root)"""
    lde = np.stack([orc.lde(orc.ntt(trace[c], 1, 1)) for c in range(trace.shape[0])])
    leaves = orc.hash_rows(lde)
    root, _ = orc.merkle(leaves)
    return lde, leaves, root


@pytest.mark.parametrize("n_log2,W,G", [(8, 64, 1), (8, 64, 2), (12, 64, 4), (13, 16, 8), (16, 8, 2), (17, 8, 4)])
def test_wide_sharded_commit_matches_oracle(n_log2, W, G):
    import torch
    import xfg_stark_b200 as xs
    from test_gpu_stages import big_ctx, rand_elems
    ctx = big_ctx()
    rng = np.random.default_rng(n_log2 * 100 + W + G)
    n = 1 << n_log2
    trace = rand_elems(rng, (W, n))
    ranks = [xs.WideTrace(ctx, n_log2, W, G, r) for r in range(G)]
    try:
        ptrs = [w.recv_ptr() for w in ranks]
        for w in ranks:
            w.set_peer_ptrs(ptrs)
        wl = W // G
        dev = [torch.from_numpy(trace[r * wl:(r + 1) * wl].view(np.int64).copy()).cuda() for r in range(G)]
        for r, w in enumerate(ranks):
            w.extend(dev[r].data_ptr())
        torch.cuda.synchronize()                       # = the cross-rank barrier of the multi-process path
        lde, leaves, root = oracle_wide(trace)
        n_local = n // G
        for r, w in enumerate(ranks):                  # receive buffer = own rows of ALL columns, coset-major
            got = w.read_recv()
            exp = lde.reshape(W, n, 8)[:, r * n_local:(r + 1) * n_local, :].transpose(0, 2, 1)
            assert (got == exp).all()
        roots = [w.commit()[0] for w in ranks]
        final = roots[0] if G == 1 else ctx.merkle_root(np.frombuffer(b"".join(roots), dtype=np.uint8).reshape(G, 32))
        assert final == root
    finally:
        for w in ranks:
            w.close()


def test_wide_two_processes_ipc():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "wide_worker.py"), "--n-log2", "14", "--cols", "64", "--check"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "WIDE_OK" in out.stdout
