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


class ActionGather:
    """In-place all-gather of the actions: one persistent [world * n_max, 2] buffer per rank; the
    solve kernel writes this rank's actions straight into its slice (`BatchedPureMPC.bind_actions(g.local)`),
    and `gather()` runs the collective with the slice as send buffer and the whole buffer as receive buffer
    (NCCL's in-place form: no staging copy, no per-step allocation).  `full` is the [total, 2] result in env
    order; with uneven shards (total % world != 0) the padding rows are skipped by a gather of row indices."""

    def __init__(self, total: int, device, dtype=torch.float32):
        init = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size() if init else 1
        self.rank = dist.get_rank() if init else 0
        self.total = int(total)
        sizes = [shard_range(total, r, self.world) for r in range(self.world)]
        self.lo, self.hi = sizes[self.rank]
        self.n_max = max(hi - lo for lo, hi in sizes)
        self.buffer = torch.zeros(self.world * self.n_max, 2, dtype=dtype, device=device)
        self.local = self.buffer[self.rank * self.n_max: self.rank * self.n_max + (self.hi - self.lo)]
        self._send = self.buffer[self.rank * self.n_max: (self.rank + 1) * self.n_max]
        self._even = all(hi - lo == self.n_max for lo, hi in sizes)
        if not self._even:
            rows = [r * self.n_max + k for r, (lo, hi) in enumerate(sizes) for k in range(hi - lo)]
            self._rows = torch.tensor(rows, dtype=torch.int64, device=device)

    def gather(self) -> torch.Tensor:
        if self.world > 1:
            if self.buffer.is_cuda:
                dist.all_gather_into_tensor(self.buffer, self._send)
            else:                                            # gloo (CPU tests): list form on views of the same buffer
                dist.all_gather([self.buffer[r * self.n_max:(r + 1) * self.n_max] for r in range(self.world)], self._send.clone())
        return self.buffer[: self.total] if self._even else self.buffer.index_select(0, self._rows)


class PipelinedActionGather:
    """Takes the all-gather off the critical path: two `ActionGather` buffers used alternately.  The collective of step
    i is issued on a side stream as soon as the solve of step i has been enqueued and runs while `k_prepare` / `k_solve`
    of step i + 1 execute; the caller only waits for it when it needs the gathered actions (or the buffer again).

        g = pipe.acquire(i)            # buffer of step i (waits for the gather of step i - 2 that used it)
        agent.bind_actions(g.local)    # the solve kernel writes this rank's slice
        agent.predict_batch(...)
        pipe.issue(i)                  # all-gather of step i on the side stream
        ...
        full = pipe.result(i)          # [total, 2]; the current stream waits for that gather

    On CPU tensors (gloo, tests) there are no streams: `issue` runs the collective synchronously."""

    def __init__(self, total: int, device, dtype=torch.float32, depth: int = 2):
        self.slots = [ActionGather(total, device, dtype) for _ in range(depth)]
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self._full = [None] * depth
        if self.cuda:
            self.side = torch.cuda.Stream(device=self.device)
            self._ready = [torch.cuda.Event() for _ in range(depth)]      # solve of the slot's step enqueued
            self._done = [None] * depth                                   # gather of the slot's step finished

    def acquire(self, i: int) -> ActionGather:
        k = i % len(self.slots)
        if self.cuda and self._done[k] is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done[k])
        return self.slots[k]

    def issue(self, i: int) -> None:
        k = i % len(self.slots)
        if not self.cuda:
            self._full[k] = self.slots[k].gather()
            return
        self._ready[k].record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.side):
            self.side.wait_event(self._ready[k])
            self._full[k] = self.slots[k].gather()
            ev = torch.cuda.Event()
            ev.record(self.side)
            self._done[k] = ev

    def result(self, i: int) -> torch.Tensor:
        k = i % len(self.slots)
        if self.cuda and self._done[k] is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done[k])
        return self._full[k]

    def drain(self) -> None:
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
