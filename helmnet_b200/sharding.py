"""Batch sharding across the GPUs of one box.

Samples are independent (no cross-sample operation anywhere in IterativeSolver.single_step, reference
helmnet/hybridnet.py:558-584), so each rank solves a contiguous slice of the batch with its own context and CUDA
graph; there is NO collective inside the loop.  After the loop the final wavefields and the residual-RMSE
histories are gathered on rank 0 (NCCL on GPUs; the same code runs on gloo for the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_slice(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the batch owned by `rank`; remainders go to the first ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(wavefield: torch.Tensor, rmse: torch.Tensor, dst: int = 0) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """Gather per-rank final wavefields [b_r,2,N,N] and RMSE histories [K,b_r] on `dst`.

    Returns (wavefield [B,2,N,N], rmse [K,B]) on dst and None elsewhere.  Ranks may own different slice sizes.
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return wavefield, rmse
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [torch.zeros(1, dtype=torch.int64, device=wavefield.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([wavefield.shape[0]], dtype=torch.int64, device=wavefield.device))
    sizes = [int(s.item()) for s in sizes]
    bmax = max(sizes)
    wf_pad = torch.zeros((bmax,) + tuple(wavefield.shape[1:]), dtype=wavefield.dtype, device=wavefield.device)
    wf_pad[: wavefield.shape[0]] = wavefield
    rm_pad = torch.zeros((rmse.shape[0], bmax), dtype=rmse.dtype, device=rmse.device)
    rm_pad[:, : rmse.shape[1]] = rmse
    wl: Optional[List[torch.Tensor]] = [torch.empty_like(wf_pad) for _ in range(world)] if rank == dst else None
    rl: Optional[List[torch.Tensor]] = [torch.empty_like(rm_pad) for _ in range(world)] if rank == dst else None
    dist.gather(wf_pad, wl, dst=dst)
    dist.gather(rm_pad, rl, dst=dst)
    if rank != dst:
        return None
    return (torch.cat([w[:s] for w, s in zip(wl, sizes)], 0), torch.cat([r[:, :s] for r, s in zip(rl, sizes)], 1))
