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


@pytest.mark.parametrize('world', [1, 2, 4, 8])
def test_sharded_norm_adj_rows_partition_the_full_matrix(world):
    """SURVEY 8 E2 host logic: every rank's CSR rows are the full L's rows (same values, same per-row order, padded
    column ids), padding rows are empty, the nonzeros partition exactly, and the overlapped nodes are a local prefix."""
    import numpy as np
    from recbole_cdr_b200.graph import NormAdj
    from recbole_cdr_b200.shard_graph import NodeShards, ShardedNormAdj
    rng = np.random.RandomState(0)
    nu, ni, ovu, ovi = 37, 23, 11, 5
    r, c = rng.randint(0, nu, 300), rng.randint(0, ni, 300)
    full = NormAdj(r, c, nu, ni, 'cpu', chunk=4)
    seen = 0
    for rank in range(world):
        nd = NodeShards(nu, ni, ovu, ovi, rank, world)
        a = ShardedNormAdj(r, c, nd, 'cpu', chunk=4)
        nodes = nd.local_nodes()
        assert nodes.size == nd.rows == a.n_rows
        for l, v in enumerate(nodes):
            b, e = a.rowptr[l].item(), a.rowptr[l + 1].item()
            if v < 0:
                assert b == e
                continue
            fb, fe = full.rowptr[v].item(), full.rowptr[v + 1].item()
            assert torch.equal(a.val[b:e], full.val[fb:fe])
            assert np.array_equal(a.col[b:e].numpy(), nd.pad(full.col[fb:fe].numpy()))
            seen += e - b
        assert all((0 <= v < ovu) == (l < nd.ov_users) for l, v in enumerate(nodes[:nd.ru]))
        assert all((v >= 0 and v - nu < ovi) == (l < nd.ov_items) for l, v in enumerate(nodes[nd.ru:]))
        # work items cover every local row's nonzeros exactly once, in pieces of <= chunk
        covered = (a.work_end - a.work_beg).sum().item()
        assert covered == a.nnz and (a.work_end - a.work_beg).max().item() <= 4
    assert seen == full.nnz
