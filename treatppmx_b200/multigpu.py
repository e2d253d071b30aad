"""Multi-GPU plumbing for the path: shard replicate datasets (and their chains) across ranks,
run independently, gather the per-dataset posterior summaries at the end.

The path shards by independent units: chains never communicate, so there is no
data-path collective — only this final gather.  One process per GPU; `torch.distributed` is the
plumbing (backend "nccl" over NVLink on a GPU box, "gloo" in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `n_items` datasets owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_chains(chain_dataset, world: int, rank: int, n_datasets: int):
    """Datasets [lo, hi) go to `rank`; returns (lo, hi, indices of the chains that follow them,
    their dataset index re-based to the shard)."""
    lo, hi = shard_range(n_datasets, world, rank)
    cd = np.asarray(chain_dataset)
    idx = np.nonzero((cd >= lo) & (cd < hi))[0]
    return lo, hi, idx, (cd[idx] - lo).astype(np.int32)


def gather_summaries(local: dict, n_datasets_total: int, group=None, dst=None):
    """All-gathers per-dataset summaries sharded with shard_range.

    local: name -> torch tensor whose leading dimension is the rank's number of datasets (CUDA
    tensors for nccl, CPU tensors for gloo; all ranks pass the same names / trailing shapes).
    Returns name -> tensor with leading dimension n_datasets_total, in dataset order.
    Shards are padded to the largest shard so one all_gather_into_tensor per array suffices."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = {}
    if world == 1:
        return {k: v.clone() for k, v in local.items()}
    sizes = [shard_range(n_datasets_total, world, r) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    for name in sorted(local):
        t = local[name].contiguous()
        pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        buf = torch.empty((world * nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, pad, group=group)
        parts = [buf[r * nmax: r * nmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
        out[name] = torch.cat(parts, dim=0)
    return out


def device_summaries_as_torch(batch, util_w=None):
    """Zero-copy torch views of a Batch's device summaries (for the NCCL gather)."""
    import torch

    dev = torch.device("cuda", batch.ctx.device)
    return {k: torch.as_tensor(v, device=dev) for k, v in batch.summaries_device(util_w).items()}
