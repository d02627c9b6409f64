"""Multi-GPU: the problem batch shards by environment index, one process per GPU.

Every MPC problem is independent; there is no exchange step inside a solve.  The
only collective is the all-gather of the actions (and, for an RL update, observations) that the
caller of the reference's `collect_rollouts` needs on every rank (agents/a2c_mpc.py:145-170).
Backend: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of the env index range: rank r owns [lo, hi).  Remainders go to the low ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def all_gather_actions(local_actions: torch.Tensor, total: int) -> torch.Tensor:
    """Gathers per-rank action blocks [n_r, 2] into the global [total, 2] tensor in env order.
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        if local_actions.shape[0] != total:
            raise ValueError("single-process gather expects the full batch")
        return local_actions
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(nmax, *local_actions.shape[1:], dtype=local_actions.dtype, device=local_actions.device)
    pad[: local_actions.shape[0]] = local_actions
    out = torch.empty(world * nmax, *local_actions.shape[1:], dtype=local_actions.dtype, device=local_actions.device)
    dist.all_gather_into_tensor(out, pad) if hasattr(dist, "all_gather_into_tensor") and local_actions.is_cuda else \
        _gather_list(out, pad, world, nmax)
    parts = [out[r * nmax: r * nmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def _gather_list(out: torch.Tensor, pad: torch.Tensor, world: int, nmax: int) -> None:
    chunks = [out[r * nmax:(r + 1) * nmax] for r in range(world)]
    dist.all_gather(chunks, pad)
