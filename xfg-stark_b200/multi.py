"""Multi-GPU plumbing for the burn-mint prover: one process per GPU, independent proofs (replicas), no collective on the data
path ("1024 independent 2^16 proofs: proof i -> GPU i mod G").  torch.distributed is used only for the
rendezvous, the barrier around the timed region and the max-over-ranks of the elapsed time; `backend` is "nccl" on the GPU
box and "gloo" in the CPU test-suite."""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def proof_indices_for_rank(total, rank, world):
    """Proof i of a batch is proven by rank i mod world (replaces the sequential loops of
    examples/winterfell_burn_mint_production.rs:187-195 and src/burn_mint_verifier.rs:326-338)."""
    return list(range(rank, total, world))


def init(backend, device=None):
    import torch.distributed as dist
    _, world, _ = rank_world()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if device is not None and backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return world


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """max of a per-rank scalar (elapsed ms of the timed region) over all ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def finalize():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
