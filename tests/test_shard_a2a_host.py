"""The all-to-all form of the row-sharded step (recbole_cdr_b200/shard_a2a.py) on CPU: two gloo processes, the kernels
(gather, fused score + loss, scatter-add -- all hardware-validated) running under the CTA emulator.  Per-rank losses must
equal the oracle's per-batch loss on that rank's batch; the re-assembled gradient tables must equal the dense gradients over
the union of the batches.  Also world size 1 in-process (the exchange degenerates to copies)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _case(world, pairwise, nu=301, ni=403, dim=32, B=48):
    g = torch.Generator().manual_seed(3)
    ut, it = torch.randn(nu, dim, generator=g) * 0.3, torch.randn(ni, dim, generator=g) * 0.3
    batches = []
    for r in range(world):
        gr = torch.Generator().manual_seed(100 + r)
        batches.append((torch.randint(0, nu, (B,), generator=gr), torch.randint(0, ni, (B,), generator=gr),
                        torch.randint(0, ni, (B,), generator=gr), (torch.rand(B, generator=gr) < 0.5).float()))
    return ut, it, batches


def _reference(ut, it, batches, pairwise, reg):
    from oracle import cdr_oracle as O
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    losses = []
    for u, ia, ib, y in batches:
        if pairwise:
            losses.append(O.emcdr_bpr_loss(a, b, u, ia, ib, reg))
        else:
            losses.append(O.bce_loss(torch.sigmoid(O.dot_score(a, b, u, ia)), y) + reg * O.emb_loss(a[u], b[ia]))
    gu, gi = torch.autograd.grad(sum(l.sum() for l in losses), [a, b])
    return [float(l.detach().reshape(-1)[0]) for l in losses], gu, gi


def _run_rank(rank, world, pairwise, scale=1.0):
    import emu_util
    from recbole_cdr_b200 import _lib
    from recbole_cdr_b200.shard import RowShardedTable
    from recbole_cdr_b200.shard_a2a import AllToAllStep
    ut, it, batches = _case(world, pairwise)
    reg = 0.02
    with emu_util.patched_ops(sms=2):
        tu, ti = RowShardedTable.from_full(ut, rank, world, 'cpu'), RowShardedTable.from_full(it, rank, world, 'cpu')
        du, di = RowShardedTable(ut.shape[0], ut.shape[1], rank, world, 'cpu'), RowShardedTable(it.shape[0], it.shape[1], rank, world, 'cpu')
        step = AllToAllStep(tu, ti, du, di, pairwise=pairwise, loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=reg)
        u, ia, ib, y = batches[rank]
        loss = step.step(u, ia, ib if pairwise else None, None if pairwise else y, scale=scale)
        assert step.exchanged_rows == u.numel() * (3 if pairwise else 2)
    ref_losses, gu, gi = _reference(ut, it, batches, pairwise, reg)
    assert abs(float(loss.reshape(-1)[0]) - ref_losses[rank]) <= 1e-4 * abs(ref_losses[rank])
    got_u, got_i = du.to_full(), di.to_full()
    torch.testing.assert_close(got_u, gu * scale, rtol=1e-4, atol=1e-4 * float(gu.abs().max()))
    torch.testing.assert_close(got_i, gi * scale, rtol=1e-4, atol=1e-4 * float(gi.abs().max()))


def _worker(rank, world, port, tmp, pairwise):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        _run_rank(rank, world, pairwise, scale=-0.5)
        open(os.path.join(tmp, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('pairwise', [True, False])
def test_all_to_all_step_world1(pairwise):
    _run_rank(0, 1, pairwise)


@pytest.mark.parametrize('pairwise', [True, False])
def test_all_to_all_step_gloo_world2(tmp_path, pairwise):
    world = 2
    port = 29700 + (os.getpid() % 2000) + (1 if pairwise else 0)
    mp.spawn(_worker, args=(world, port, str(tmp_path), pairwise), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))


# ---- K steps per exchange (AllToAllChunkRunner): the persistent kernel on block-sized mini tables --------------------------------

def _run_rank_chunk(rank, world, pairwise, K=3, B=32):
    import emu_util
    from recbole_cdr_b200 import _lib
    from recbole_cdr_b200.shard import RowShardedTable
    from recbole_cdr_b200.shard_a2a import AllToAllChunkRunner
    from oracle import cdr_oracle as O
    nu, ni, dim, reg = 203, 301, 32, 0.02
    g = torch.Generator().manual_seed(5)
    ut, it = torch.randn(nu, dim, generator=g) * 0.3, torch.randn(ni, dim, generator=g) * 0.3
    blocks = []
    for r in range(world):
        gr = torch.Generator().manual_seed(200 + r)
        ids = torch.stack([torch.randint(0, nu, (K, B), generator=gr), torch.randint(0, ni, (K, B), generator=gr)] +
                          ([torch.randint(0, ni, (K, B), generator=gr)] if pairwise else []), dim=1)
        blocks.append((ids, (torch.rand(K, B, generator=gr) < 0.5).float()))
    with emu_util.patched_ops(sms=2):
        tu, ti = RowShardedTable.from_full(ut, rank, world, 'cpu'), RowShardedTable.from_full(it, rank, world, 'cpu')
        du, di = RowShardedTable(nu, dim, rank, world, 'cpu'), RowShardedTable(ni, dim, rank, world, 'cpu')
        runner = AllToAllChunkRunner(tu, ti, du, di, pairwise=pairwise, loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=reg)
        ids, y = blocks[rank]
        lab = None
        if not pairwise:     # the persistent kernel wants the label rows on the id block's step stride
            lab = torch.zeros((K, 2, B))[:, 0]
            lab.copy_(y)
        losses = runner.run(ids, lab)
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    total = 0
    for r in range(world):
        ids_r, y_r = blocks[r]
        for k in range(K):
            if pairwise:
                ref = O.emcdr_bpr_loss(a, b, ids_r[k, 0], ids_r[k, 1], ids_r[k, 2], reg)
            else:
                ref = O.bce_loss(torch.sigmoid(O.dot_score(a, b, ids_r[k, 0], ids_r[k, 1])), y_r[k]) + reg * O.emb_loss(a[ids_r[k, 0]], b[ids_r[k, 1]])
            if r == rank:
                assert abs(float(losses[k]) - float(ref.detach().reshape(-1)[0])) <= 1e-4 * abs(float(ref.detach().reshape(-1)[0]))
            total = total + ref.sum()
    gu, gi = torch.autograd.grad(total, [a, b])
    torch.testing.assert_close(du.to_full(), gu, rtol=1e-4, atol=1e-4 * float(gu.abs().max()))
    torch.testing.assert_close(di.to_full(), gi, rtol=1e-4, atol=1e-4 * float(gi.abs().max()))


def _chunk_worker(rank, world, port, tmp, pairwise):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        _run_rank_chunk(rank, world, pairwise)
        open(os.path.join(tmp, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('pairwise', [True, False])
def test_all_to_all_chunk_world1(pairwise):
    _run_rank_chunk(0, 1, pairwise)


@pytest.mark.parametrize('pairwise', [True, False])
def test_all_to_all_chunk_gloo_world2(tmp_path, pairwise):
    world = 2
    port = 30100 + (os.getpid() % 2000) + (1 if pairwise else 0)
    mp.spawn(_chunk_worker, args=(world, port, str(tmp_path), pairwise), nprocs=world, join=True)
    assert all((tmp_path / f'ok{r}').exists() for r in range(world))
