"""Data-parallel frame sharding: frames are independent units, rank r owns a contiguous block,
weights are replicated, inference needs NO data-path collective.  The helpers below are the only
`torch.distributed` use outside bench.py: a max-over-ranks for timings and a gather used to CHECK that the
N-rank result equals the 1-rank result."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int) -> range:
    """Contiguous block of frame ids for `rank`; the first n_frames % world ranks get one extra frame."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def max_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(rows: torch.Tensor, dst: int = 0) -> List[torch.Tensor]:
    """Gather variable-length (n_i, k) row blocks on `dst` (checking only -- not on the timed path)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [rows]
    world = dist.get_world_size()
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    cap = int(max(int(c.item()) for c in counts))
    buf = torch.zeros((cap,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    buf[: rows.shape[0]] = rows
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return [o[: int(c.item())] for o, c in zip(out, counts)]


def assign(frames: Sequence, rank: int, world: int) -> list:
    return [frames[i] for i in shard_range(len(frames), rank, world)]


class FlatGradExchange:
    """The training path's one exchange step: a sum over the flat fp32 gradient buffer, issued as two
    collectives so that the tail of the buffer (the fusion head: fc6/fc7, finished first in backward) travels while the
    trunks are still being differentiated.  Works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self._pending = None

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def start_tail(self, buf: torch.Tensor, off: int) -> None:
        if self.world > 1 and off < buf.numel():
            self._pending = dist.all_reduce(buf[off:], group=self.group, async_op=True)

    def finish(self, buf: torch.Tensor, off: int) -> float:
        """Reduce the head of the buffer, wait for the tail; returns the 1/world factor the optimizer applies."""
        if self.world > 1:
            if off > 0:
                dist.all_reduce(buf[:off], group=self.group)
            if self._pending is not None:
                self._pending.wait()
                self._pending = None
        return 1.0 / self.world
