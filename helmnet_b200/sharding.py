"""Batch sharding across the GPUs of one box.

Samples are independent (no cross-sample operation anywhere in IterativeSolver.single_step, reference
helmnet/hybridnet.py:558-584), so each rank solves a contiguous slice of the batch with its own context and CUDA
graph; there is NO collective inside the loop.  After the loop the final wavefields and the residual-RMSE
histories are gathered on rank 0 (NCCL on GPUs; the same code runs on gloo for the CPU tests).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_slice(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the batch owned by `rank`; remainders go to the first ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(batch: int, world: int) -> List[int]:
    return [shard_slice(batch, world, r)[1] - shard_slice(batch, world, r)[0] for r in range(world)]


def gather_results(wavefield: torch.Tensor, rmse: torch.Tensor, dst: int = 0,
                   sizes: Optional[Sequence[int]] = None) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """Gather per-rank final wavefields [b_r,2,N,N] and RMSE histories [K,b_r] on `dst`.

    Returns (wavefield [B,2,N,N], rmse [K,B]) on dst and None elsewhere.  ONE collective: every rank packs its wavefield
    slice and its RMSE history into one contiguous float buffer (padded to the largest slice when the slices differ) and
    `dst` receives them into one preallocated [world, len] tensor.  `sizes` (the per-rank slice sizes, e.g.
    `shard_sizes(batch, world)`) saves the size exchange; without it the sizes are all-gathered first.
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return wavefield, rmse
    world, rank = dist.get_world_size(), dist.get_rank()
    if sizes is None:
        cells = [torch.zeros(1, dtype=torch.int64, device=wavefield.device) for _ in range(world)]
        dist.all_gather(cells, torch.tensor([wavefield.shape[0]], dtype=torch.int64, device=wavefield.device))
        sizes = [int(s.item()) for s in cells]
    sizes = [int(s) for s in sizes]
    if len(sizes) != world or sizes[rank] != wavefield.shape[0] or rmse.shape[1] != wavefield.shape[0]:
        raise ValueError("slice sizes do not match this rank's tensors")
    bmax, k = max(sizes), rmse.shape[0]
    per_wf = math.prod(wavefield.shape[1:])
    n_wf, n_rm = bmax * per_wf, k * bmax
    buf = torch.zeros(n_wf + n_rm, dtype=wavefield.dtype, device=wavefield.device) if sizes[rank] != bmax else \
        torch.empty(n_wf + n_rm, dtype=wavefield.dtype, device=wavefield.device)
    buf[: wavefield.numel()].copy_(wavefield.reshape(-1))
    buf[n_wf:].view(k, bmax)[:, : sizes[rank]].copy_(rmse)
    if rank == dst:
        recv = torch.empty(world, n_wf + n_rm, dtype=wavefield.dtype, device=wavefield.device)
        dist.gather(buf, [recv[r] for r in range(world)], dst=dst)
    else:
        dist.gather(buf, None, dst=dst)
        return None
    shape = tuple(wavefield.shape[1:])
    if all(s == bmax for s in sizes):
        wf = recv[:, :n_wf].reshape((world * bmax,) + shape)
        rm = recv[:, n_wf:].reshape(world, k, bmax).permute(1, 0, 2).reshape(k, world * bmax)
        return wf, rm
    wf = torch.cat([recv[r, :n_wf].view((bmax,) + shape)[: sizes[r]] for r in range(world)], 0)
    rm = torch.cat([recv[r, n_wf:].view(k, bmax)[:, : sizes[r]] for r in range(world)], 1)
    return wf, rm
