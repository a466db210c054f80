"""BiTGCF training step over 1..G GPUs (BASELINE config #4 shape scaled by --scale).

  python scripts/bench_bitgcf.py                      # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_bitgcf.py

Full config #4 = 2M users x 1M items per domain, 50 % user overlap, 16 edges per user (32M edges per domain), 3 layers,
D = 64, connect_way mean, B = 16384 per domain.  --scale 0.25 (default) keeps the host graph build short.
Prints one JSON line (rank 0): ms per step (CUDA events, max over ranks) and edges*layers/s
(= 2 domains * E * n_layers * 2 passes (fwd + bwd SpMM) / step time counted as in SURVEY 8 D3: nnz_A = 2E per SpMM)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'recbole-cdr_b200'))
from recbole_cdr_b200.shard_graph import ShardedBiTGCF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--scale', type=float, default=0.25)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--layers', type=int, default=3)
ap.add_argument('--dim', type=int, default=64)
ap.add_argument('--batch', type=int, default=16384)
ap.add_argument('--exchange', default='allgather', choices=['allgather', 'peer'])
args = ap.parse_args()

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)

# joint id layout: [0, n_ov) overlapped users, then target-only, then source-only; items disjoint
per_dom_users, per_dom_items = int(2_000_000 * args.scale), int(1_000_000 * args.scale)
n_ov = per_dom_users // 2 + 1
n_only = per_dom_users - (n_ov - 1)
nu, ni = n_ov + 2 * n_only, 1 + 2 * per_dom_items
t0 = time.time()
def edges(seed, user_ids, item_lo):
    rng = np.random.RandomState(seed)
    u = np.repeat(user_ids, 16)
    i = item_lo + np.minimum(rng.zipf(1.05, u.size) - 1, per_dom_items - 1)
    return u, i
tgt_users = np.arange(1, n_ov + n_only)
src_users = np.concatenate([np.arange(1, n_ov), np.arange(n_ov + n_only, nu)])
src, tgt = edges(0, src_users, 1 + per_dom_items), edges(1, tgt_users, 1)
eng = ShardedBiTGCF(src, tgt, nu, ni, n_ov, 1, dim=args.dim, n_layers=args.layers, lambda_source=0.8, lambda_target=0.8,
                    connect_way='mean', reg_weight=0.001, rank=rank, world=world, device=dev, exchange=args.exchange)
build_s = time.time() - t0
nnz = eng.adj_s.nnz + eng.adj_t.nnz                    # this rank's nonzeros of both L
g = torch.Generator().manual_seed(100 + rank)
def batch(users, item_lo):
    u = torch.from_numpy(users)[torch.randint(0, users.size, (args.batch,), generator=g)]
    i = item_lo + torch.randint(0, per_dom_items, (args.batch,), generator=g)
    y = (torch.rand(args.batch, generator=g) < 0.5).float()
    return u.to(dev), i.to(dev), y.to(dev)
bs, bt = batch(src_users, 1 + per_dom_items), batch(tgt_users, 1)

def step():
    eng.ego_s.local.grad = eng.ego_t.local.grad = None
    return eng.train_step(bs, bt)

for _ in range(args.warmup):
    losses = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    losses = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
tot = torch.tensor([float(nnz)], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(tot)
if rank == 0:
    N = nu + ni
    nnz_all = tot.item()                                   # = 2E (source) + 2E (target)
    # SURVEY 8 D3 per layer-domain pass: nnz*(12 + 4D) + 2*N*4D; passes = layers * 2 (fwd + bwd), both domains in nnz_all
    spmm_bytes = args.layers * 2 * (nnz_all * (12 + 4 * args.dim) + 2 * 2 * N * 4 * args.dim)
    print(json.dumps({'bench': 'bitgcf_train_step', 'n_gpus': world, 'exchange': eng.exchange, 'scale': args.scale, 'users': nu, 'items': ni,
                      'nnz_L_both_domains': nnz_all, 'layers': args.layers, 'dim': args.dim, 'batch_per_domain_per_gpu': args.batch,
                      'ms_per_step': ms.item(), 'edges_layers_per_s': (nnz_all / 2) * args.layers / (ms.item() * 1e-3),
                      'spmm_algorithmic_GBps_aggregate': spmm_bytes / (ms.item() * 1e-3) / 1e9,
                      'loss': [float(l) for l in losses], 'host_graph_build_s': build_s}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
