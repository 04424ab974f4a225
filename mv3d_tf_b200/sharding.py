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
    """The training path's one exchange step: a sum over the flat fp32 gradient buffer, issued as
    BUCKETS while backward is still running.  Backward finalises the buffer from its tail to its head (fusion head
    first, then the RPN, then the trunks last layer first), so `ready(buf, lo)` is called with a falling low-water
    mark; whenever at least `bucket_bytes` of finished gradients have accumulated, that contiguous slice goes out as
    one asynchronous all-reduce (NCCL: on the communicator's own stream, ordered after the producing kernels).  Nothing
    blocks until `finish`, which flushes the remainder and makes the current stream wait for every bucket.  Works on
    any backend (NCCL on the GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None, bucket_bytes: int = 32 << 20):
        self.group = group
        self.bucket_bytes = int(bucket_bytes)
        self._pending = []
        self._hi = {}            # region key -> elements [hi, region end) of that region are already on their way
        self._regions = {}       # region key -> (lo, hi) element bounds; key 0 = the whole buffer unless narrowed
        self.buckets_last_step = 0

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def set_regions(self, regions: dict) -> None:
        """Disjoint element ranges {key: (lo, hi)} that together cover the buffer.  Each region is finalised from its top
        to its bottom by ONE stream (the two trunks run their backward on two streams), with its own low-water mark."""
        self._regions = {k: (int(a), int(b)) for k, (a, b) in regions.items()}

    def _send(self, buf: torch.Tensor, lo: int, hi: int) -> None:
        if hi > lo:
            self._pending.append(dist.all_reduce(buf[lo:hi], group=self.group, async_op=True))

    def _bounds(self, buf, key):
        return self._regions.get(key, (0, buf.numel()))

    def ready(self, buf: torch.Tensor, lo: int, force: bool = False, key=0) -> None:
        """Gradients in buf[lo : end of region `key`] are final.  Call it on the stream that produced them: the
        collective is ordered after that stream's work."""
        if self.world <= 1:
            return
        rlo, rhi = self._bounds(buf, key)
        hi = self._hi.get(key, rhi)
        lo = max(rlo, min(int(lo), hi))
        if force or (hi - lo) * buf.element_size() >= self.bucket_bytes:
            self._send(buf, lo, hi)
            self._hi[key] = lo

    def start_tail(self, buf: torch.Tensor, off: int, key=0) -> None:   # first bucket: everything from `off` to the region end
        self.ready(buf, off, force=True, key=key)

    def finish(self, buf: torch.Tensor, off: int = 0) -> float:
        """Send what is left of every region, wait for every bucket; returns the 1/world factor the optimizer applies.
        Call it after the producing streams were joined into the current one."""
        if self.world > 1:
            keys = list(self._regions) if self._regions else [0]
            for key in keys:
                rlo, rhi = self._bounds(buf, key)
                self._send(buf, rlo, self._hi.get(key, rhi))
            self.buckets_last_step = len(self._pending)
            for h in self._pending:
                h.wait()
            self._pending = []
            self._hi = {}
        return 1.0 / self.world
