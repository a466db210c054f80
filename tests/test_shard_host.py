"""CPU tests (gloo, world_size 2) of the host-side logic of the row-sharded path: layout math, shard <-> full round trip
through a real process group, and user-owner batch routing."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_layout_math():
    from recbole_cdr_b200 import shard
    rows = torch.arange(0, 1001)
    for world in (1, 2, 4, 8):
        own, loc = shard.owner_of(rows, world), shard.local_row_of(rows, world)
        assert torch.equal(loc * world + own, rows)                      # bijection global <-> (owner, local)
        assert int(loc.max()) < shard.shard_rows(1001, world)
        counts = torch.bincount(own, minlength=world)
        assert int(counts.max() - counts.min()) <= 1                     # block-cyclic: perfectly balanced
    # the overlapped id range [1, n_ov) of the joint layout is spread over all ranks (a block split would not)
    assert set(shard.owner_of(torch.arange(1, 9), 4).tolist()) == {0, 1, 2, 3}
    with pytest.raises(ValueError):
        shard.RowShardedTable(10, 4, 0, 3, 'cpu')                        # world must be a power of two


def test_user_owner_routing_partitions_the_batch():
    from recbole_cdr_b200 import shard
    g = torch.Generator().manual_seed(0)
    user = torch.randint(1, 10_000, (4096,), generator=g)
    parts = shard.route_by_user_owner(user, 4)
    assert sum(p.numel() for p in parts) == user.numel()
    assert torch.equal(torch.sort(torch.cat(parts)).values, torch.arange(user.numel()))
    for r, p in enumerate(parts):
        assert (user[p] % 4 == r).all()


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from recbole_cdr_b200 import shard
        g = torch.Generator().manual_seed(7)
        full = torch.randn(1001, 8, generator=g)          # identical on every rank (same seed)
        t = shard.RowShardedTable.from_full(full, rank, world, 'cpu')
        assert t.local.shape == (shard.shard_rows(1001, world), 8)
        assert torch.equal(t.local[:full[rank::world].shape[0]], full[rank::world])
        back = t.to_full()
        assert torch.equal(back, full)
        # data-parallel bookkeeping: per-rank step losses combine to the global mean by batch size
        losses = torch.tensor([float(rank + 1)])
        dist.all_reduce(losses)
        assert losses.item() == sum(range(1, world + 1))
        open(os.path.join(tmp, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_shard_round_trip_gloo_world2(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))
