"""N>1 host logic on CPU: two gloo ranks shard frames with no data-path collective, and the gathered result
equals the single-rank result (frame order and content)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_detect(frame_id: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 + frame_id)
    n = 3 + frame_id % 4
    return torch.cat([torch.full((n, 1), float(frame_id)), torch.rand((n, 5), generator=g)], dim=1)


def _worker(rank, world, port, n_frames, q):
    from mv3d_tf_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_range(n_frames, rank, world)
    rows = torch.cat([_fake_detect(f) for f in mine]) if len(mine) else torch.zeros((0, 6))
    parts = sharding.gather_rows(rows)
    t = sharding.max_over_ranks(1.0 + rank)
    if rank == 0:
        q.put((torch.cat(parts), t, [list(sharding.shard_range(n_frames, r, world)) for r in range(world)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [64, 7, 1])
def test_two_rank_sharding_matches_single_rank(n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_frames) % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, tmax, shards = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.cat([_fake_detect(f) for f in range(n_frames)])
    assert torch.equal(got, want)
    assert tmax == 2.0
    assert sorted(sum(shards, [])) == list(range(n_frames))


def test_shard_range_properties():
    from mv3d_tf_b200.sharding import shard_range

    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            parts = [list(shard_range(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _exchange_worker(rank, world, port, q):
    from mv3d_tf_b200.sharding import FlatGradExchange

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = FlatGradExchange(dist.group.WORLD)
    n, off = 1000, 300
    buf = torch.arange(n, dtype=torch.float32) * (rank + 1)
    ex.start_tail(buf, off)                  # the head's gradients travel first ...
    buf[:off] += 0.5                         # ... while the "trunk backward" still writes its part
    scale = ex.finish(buf, off)
    # bucketed form: backward hands over a falling low-water mark; slices go out whenever >= bucket_bytes are final
    ex2 = FlatGradExchange(dist.group.WORLD, bucket_bytes=4 * 128)
    b2 = torch.arange(n, dtype=torch.float32) * (rank + 1)
    marks = [900, 880, 700, 690, 400, 399, 120, 0]
    for lo in marks:
        ex2.ready(b2, lo)
    ex2.finish(b2)
    if rank == 0:
        q.put((buf.clone(), scale, b2.clone(), ex2.buckets_last_step))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_exchange_two_ranks():
    """Training exchange step: sum over ranks of the flat buffer in two collectives; the optimizer's 1/world factor."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 77) % 500
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    buf, scale, b2, n_buckets = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.arange(1000, dtype=torch.float32) * 3
    want[:300] += 1.0
    assert torch.equal(buf, want) and scale == 0.5
    assert torch.equal(b2, torch.arange(1000, dtype=torch.float32) * 3)   # every element reduced exactly once
    assert 3 <= n_buckets <= 6
