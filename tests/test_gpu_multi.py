"""Multi-GPU parity of the row-sharded persistent path (needs >= 2 GPUs: run under `gpurun --gpus 2`).

G ranks, block-cyclic shards mapped over CUDA IPC, each rank trains on its own batches; the per-batch losses and the
accumulated gradients must equal a single-process run of the per-step kernels over the union of the batches."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp, staged):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from recbole_cdr_b200 import ops, shard
        nu, ni, dim, K, B = 4001, 5003, 64, 6, 2048
        g = torch.Generator().manual_seed(3)
        ut, it = torch.randn(nu, dim, generator=g) * 0.1, torch.randn(ni, dim, generator=g) * 0.1
        # every rank draws ALL ranks' batches (same seed) so that rank 0 can run the single-GPU reference
        u = torch.randint(0, nu // world, (world, K, B), generator=g) * world + torch.arange(world).view(-1, 1, 1)
        u = u.clamp_max(nu - 1)                                     # user-owner routing: u % world == rank (mostly)
        ip = torch.randint(0, ni, (world, K, B), generator=g)
        ineg = torch.randint(0, ni, (world, K, B), generator=g)
        tabs = [shard.RowShardedTable.from_full(t, rank, world, dev).connect() for t in (ut, it)]
        grads = [shard.RowShardedTable(t.shape[0], dim, rank, world, dev).connect() for t in (ut, it)]
        dist.barrier()
        if staged:   # peer-gather kernel one chunk ahead + dense staged item rows (chunks of 4 steps, ragged last chunk)
            runner = shard.ShardedStepRunner(tabs[0], tabs[1], grads[0], grads[1], reg_weight=0.01, chunk=4, stage_remote=True)
            ids = torch.stack([u[rank], ip[rank], ineg[rank]], dim=1).to(dev)
            out8 = runner.run(ids)
        else:        # item rows gathered straight from the peer shards inside the persistent kernel
            out8 = shard.train_steps_sharded(tabs[0], tabs[1], grads[0], grads[1], u[rank].to(dev), ip[rank].to(dev),
                                             ineg[rank].to(dev), reg_weight=0.01)
        torch.cuda.synchronize()
        dist.barrier()                                              # every rank's remote REDs have landed
        gu_full, gi_full = grads[0].to_full(), grads[1].to_full()
        losses = [torch.empty_like(out8[:, 0]) for _ in range(world)]
        dist.all_gather(losses, out8[:, 0].contiguous())
        if rank == 0:
            a, b = ut.to(dev).requires_grad_(True), it.to(dev).requires_grad_(True)
            for r in range(world):
                for k in range(K):
                    loss = ops.bpr_loss(a, b, u[r, k].to(dev), ip[r, k].to(dev), ineg[r, k].to(dev), 0.01)
                    torch.testing.assert_close(losses[r][k:k + 1], loss.detach(), rtol=2e-6, atol=0)
                    loss.backward()
            torch.testing.assert_close(gu_full, a.grad, rtol=1e-4, atol=1e-8)
            torch.testing.assert_close(gi_full, b.grad, rtol=1e-4, atol=1e-8)
            open(os.path.join(tmp, 'ok'), 'w').write('ok')
        dist.barrier()
        for t in tabs + grads:
            t.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('staged', [False, True])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_steps_match_single_gpu(world, staged, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    port = 29600 + (os.getpid() % 2000) + world + (10 if staged else 0)
    mp.spawn(_worker, args=(world, port, str(tmp_path), staged), nprocs=world, join=True)
    assert (tmp_path / 'ok').exists()


# ---- E2: BiTGCF with row-sharded graph propagation ---------------------------------------------------------------------

def _bitgcf_worker(rank, world, port, tmp, way, exchange='peer'):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    if world > 1:
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        import numpy as np
        from fake_data import FakeDataset, base_config
        from recbole_cdr_b200.shard_graph import ShardedBiTGCF
        ds_args = (101, 90, 110, 31, 80, 92)                                  # overlapped users AND items, odd sizes
        nu, ni, D, L, B = 301, 203, 16, 2, 512
        rng = np.random.RandomState(5)
        edges = {'source': (rng.randint(0, nu, 3000), np.minimum(rng.zipf(1.3, 3000) - 1, ni - 1)),   # one hub item > 256 nnz
                 'target': (rng.randint(0, nu, 2500), rng.randint(0, ni, 2500))}
        g = torch.Generator().manual_seed(11)
        ego = [torch.randn(n, D, generator=g) * 0.3 for n in (nu, ni, nu, ni)]   # source_user, source_item, target_user, target_item
        bu = torch.randint(0, nu, (world, 2, B), generator=g)
        bi = torch.randint(0, ni, (world, 2, B), generator=g)
        by = (torch.rand(world, 2, B, generator=g) < 0.5).float()
        eng = ShardedBiTGCF(edges['source'], edges['target'], nu, ni, ds_args[0], ds_args[3], dim=D, n_layers=L,
                            lambda_source=0.8, lambda_target=0.7, connect_way=way, reg_weight=0.01, rank=rank, world=world,
                            device=dev, ego=[t.to(dev) for t in ego], exchange=exchange)
        mine = [(bu[rank, d].to(dev), bi[rank, d].to(dev), by[rank, d].to(dev)) for d in range(2)]
        for _ in range(2):                    # twice: the second pass re-uses every exchange buffer
            eng.ego_s.local.grad = eng.ego_t.local.grad = None
            loss_s, loss_t = eng.train_step(*mine)
        grads = eng.full_tables('grad')
        tabs = eng.full_tables('ego')
        all_losses = [torch.cat([loss_s, loss_t])] if world == 1 else [torch.empty(2, device=dev) for _ in range(world)]
        if world > 1:
            dist.all_gather(all_losses, torch.cat([loss_s, loss_t]))
        if rank == 0:
            from recbole_cdr_b200.data.interaction import Interaction
            from recbole_cdr_b200.model.cross_domain_recommender.bitgcf import BiTGCF
            m = BiTGCF(base_config(device=dev, embedding_size=D, n_layers=L, reg_weight=0.01, lambda_source=0.8, lambda_target=0.7,
                                   drop_rate=0.0, connect_way=way), FakeDataset(*ds_args, edges=edges)).to(dev)
            names = ('source_user_embedding', 'source_item_embedding', 'target_user_embedding', 'target_item_embedding')
            with torch.no_grad():
                for n, t, back in zip(names, ego, tabs):
                    getattr(m, n).weight.copy_(t)
                    assert torch.equal(back, t.to(dev))                      # shard -> full round trip is bit-exact
            for r in range(world):
                ls, lt = m.calculate_loss(Interaction({
                    'source_user_id': bu[r, 0].to(dev), 'source_item_id': bi[r, 0].to(dev), 'source_label': by[r, 0].to(dev),
                    'target_user_id': bu[r, 1].to(dev), 'target_item_id': bi[r, 1].to(dev), 'target_label': by[r, 1].to(dev)}))
                # fp32 loss within 1e-4 relative (north_star); observed ~1e-6
                torch.testing.assert_close(all_losses[r], torch.cat([ls, lt]).detach(), rtol=1e-4, atol=0)
                ((ls + lt) / world).sum().backward()
            for n, got in zip(names, grads):
                want = getattr(m, n).weight.grad
                torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4 * want.abs().max().item())
            open(os.path.join(tmp, 'ok'), 'w').write('ok')
        if world > 1:
            dist.barrier()
        eng.close()
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.parametrize('exchange', ['allgather', 'peer'])
@pytest.mark.parametrize('way', ['concat', 'mean'])
@pytest.mark.parametrize('world', [1, 2, 4])
def test_sharded_bitgcf_step_matches_single_gpu(world, way, exchange, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    port = 31700 + (os.getpid() % 2000) + world + (20 if way == 'mean' else 0) + (40 if exchange == 'peer' else 0)
    if world == 1:      # the same engine, one shard, no process group: runs on the single-GPU box too
        if exchange == 'allgather':
            pytest.skip('one shard: nothing to exchange')
        _bitgcf_worker(0, 1, port, str(tmp_path), way)
    else:
        mp.spawn(_bitgcf_worker, args=(world, port, str(tmp_path), way, exchange), nprocs=world, join=True)
    assert (tmp_path / 'ok').exists()


# ---- E1, all-to-all form (shard_a2a.py): the exchange BASELINE.json's north_star names, over NCCL -------------------------

def _a2a_worker(rank, world, port, tmp, chunked):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from recbole_cdr_b200 import ops, shard
        from recbole_cdr_b200.shard_a2a import AllToAllStep
        nu, ni, dim, K, B = 4001, 5003, 64, 3, 2048
        g = torch.Generator().manual_seed(3)
        ut, it = torch.randn(nu, dim, generator=g) * 0.1, torch.randn(ni, dim, generator=g) * 0.1
        u = torch.randint(0, nu, (world, K, B), generator=g)
        ip, ineg = torch.randint(0, ni, (world, K, B), generator=g), torch.randint(0, ni, (world, K, B), generator=g)
        tabs = [shard.RowShardedTable.from_full(t, rank, world, dev) for t in (ut, it)]
        grads = [shard.RowShardedTable(t.shape[0], dim, rank, world, dev) for t in (ut, it)]
        step = AllToAllStep(tabs[0], tabs[1], grads[0], grads[1], pairwise=True, reg_weight=0.01)
        if chunked:     # one exchange round for the K steps + the persistent kernel on block-sized mini tables
            from recbole_cdr_b200.shard_a2a import AllToAllChunkRunner
            runner = AllToAllChunkRunner(tabs[0], tabs[1], grads[0], grads[1], pairwise=True, reg_weight=0.01)
            mine = runner.run(torch.stack([u[rank], ip[rank], ineg[rank]], dim=1).contiguous().to(dev)).clone()
        else:
            mine = torch.cat([step.step(u[rank, k].to(dev), ip[rank, k].to(dev), ineg[rank, k].to(dev)).reshape(-1) for k in range(K)])
        torch.cuda.synchronize()
        dist.barrier()
        gu_full, gi_full = grads[0].to_full(), grads[1].to_full()
        losses = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(losses, mine)
        if rank == 0:
            a, b = ut.to(dev).requires_grad_(True), it.to(dev).requires_grad_(True)
            for r in range(world):
                for k in range(K):
                    loss = ops.bpr_loss(a, b, u[r, k].to(dev), ip[r, k].to(dev), ineg[r, k].to(dev), 0.01)
                    torch.testing.assert_close(losses[r][k:k + 1], loss.detach().reshape(-1), rtol=1e-5, atol=0)
                    loss.backward()
            torch.testing.assert_close(gu_full, a.grad, rtol=1e-4, atol=1e-8)
            torch.testing.assert_close(gi_full, b.grad, rtol=1e-4, atol=1e-8)
            open(os.path.join(tmp, 'ok'), 'w').write('ok')
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.unvalidated
@pytest.mark.parametrize('chunked', [False, True])
@pytest.mark.parametrize('world', [2, 4])
def test_all_to_all_steps_match_single_gpu(world, chunked, tmp_path):
    """Same equivalence as the peer-memory path, through NCCL all-to-alls (hardware-validated kernels, a host path that has
    only run over gloo so far -- hence `unvalidated`)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    port = 29900 + (os.getpid() % 2000) + world + (20 if chunked else 0)
    mp.spawn(_a2a_worker, args=(world, port, str(tmp_path), chunked), nprocs=world, join=True)
    assert (tmp_path / 'ok').exists()
