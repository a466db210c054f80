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
