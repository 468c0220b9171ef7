"""Multi-GPU sharding of a batch of hoppers: contiguous index ranges, one process per GPU.

The hot path has no exchange step (every hopper is independent), so there is no collective on
it.  The only communication is the optional gather of logged trajectories / statistics after a rollout:
``gather_hoppers`` (NCCL over NVLink for CUDA tensors, gloo for CPU tensors in the tests).
Scenarios are keyed by the GLOBAL hopper index (scenarios.make_batch(idx0=...)), so results do not depend
on the number of shards."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(B_total, rank, world):
    """Contiguous range [lo, hi) of global hopper indices owned by ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(int(B_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def gather_hoppers(t, B_total, dst=None):
    """Concatenate per-rank tensors along their LAST (hopper) axis in global hopper order.

    ``t`` has shape (..., B_rank) with B_rank = this rank's shard size.  Returns the (..., B_total) tensor
    on every rank (dst=None, all_gather) or on rank ``dst`` only (others get None)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(B_total, r, world) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(t.shape[:-1] + (bmax,), dtype=t.dtype, device=t.device)
    pad[..., :t.shape[-1]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous())
    if dst is not None and rank != dst:
        return None
    return torch.cat([p[..., :hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=-1)


def max_over_ranks(value, device):
    """Max of a Python float over ranks (timing is reported as the slowest rank's)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    v = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v.item())


def sum_over_ranks(value, device):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    v = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return float(v.item())
