"""The all-to-all form of the row-sharded EMCDR step (recbole_cdr_b200/shard_a2a.py) next to the peer-memory kernel, on the
same tables and batches: what "a single NCCL all-to-all of looked-up rows per batch" costs per step against in-kernel peer
loads / REDs.  One process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/bench_a2a.py

Three forms: one exchange round per step (AllToAllStep), one per block of --steps steps with the persistent kernel on
block-sized mini tables (AllToAllChunkRunner), and the peer-memory persistent kernel.
Prints one JSON line (rank 0): us per step of the three forms (CUDA events, max over ranks) and the aggregate interactions/s."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'recbole-cdr_b200'))
from recbole_cdr_b200 import shard  # noqa: E402
from recbole_cdr_b200.shard_a2a import AllToAllChunkRunner, AllToAllStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--rows', type=int, default=1_000_000, help='rows per table PER GPU (weak scaling, as bench.py)')
ap.add_argument('--dim', type=int, default=64)
ap.add_argument('--batch', type=int, default=8192)
ap.add_argument('--steps', type=int, default=20)
ap.add_argument('--warmup', type=int, default=3)
args = ap.parse_args()

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)

n = args.rows * world
g = torch.Generator(device=dev).manual_seed(rank)
rows = shard.shard_rows(n, world)
tabs = [shard.RowShardedTable(n, args.dim, rank, world, dev, torch.randn(rows, args.dim, device=dev, generator=g) * 0.1)
        for _ in range(2)]
grads = [shard.RowShardedTable(n, args.dim, rank, world, dev) for _ in range(2)]
K, B = args.steps, args.batch
u = torch.randint(0, n // world, (K, B), device=dev, generator=g) * world + rank      # user-owner routing, as bench.py
ip, ineg = torch.randint(0, n, (K, B), device=dev, generator=g), torch.randint(0, n, (K, B), device=dev, generator=g)


def timed(fn):
    for _ in range(args.warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


a2a = AllToAllStep(tabs[0], tabs[1], grads[0], grads[1], pairwise=True, reg_weight=0.01)
t_a2a = timed(lambda: [a2a.step(u[k], ip[k], ineg[k]) for k in range(K)]) / K

chunk = AllToAllChunkRunner(tabs[0], tabs[1], grads[0], grads[1], pairwise=True, reg_weight=0.01)
block = torch.stack([u, ip, ineg], dim=1).contiguous()          # [K, 3, B]: one exchange round for all K steps
t_chunk = timed(lambda: chunk.run(block)) / K

for t in tabs + grads:
    t.connect()
if world > 1:
    dist.barrier()
t_peer = timed(lambda: shard.train_steps_sharded(tabs[0], tabs[1], grads[0], grads[1], u, ip, ineg, reg_weight=0.01)) / K

if rank == 0:
    print(json.dumps({'n_gpus': world, 'rows_per_gpu': args.rows, 'dim': args.dim, 'batch_per_gpu': B, 'steps': K,
                      'all_to_all_us_per_step': round(t_a2a * 1e6, 1), 'all_to_all_block_us_per_step': round(t_chunk * 1e6, 1),
                      'peer_kernel_us_per_step': round(t_peer * 1e6, 1),
                      'all_to_all_Minter_per_s': round(world * B / t_a2a / 1e6, 1),
                      'all_to_all_block_Minter_per_s': round(world * B / t_chunk / 1e6, 1),
                      'peer_kernel_Minter_per_s': round(world * B / t_peer / 1e6, 1)}), flush=True)
for t in tabs + grads:
    t.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
