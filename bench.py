#!/usr/bin/env python
"""bench.py -- interactions/s through the gather -> map -> score -> scatter hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl xdr|reference] [--workload emcdr_1m|emcdr_10m]

Workload (config.workload): BASELINE.json configs[1] -- EMCDR, synthetic 1M x 1M users/items per domain, dim 64,
batch 8192 (user, item+, item-) triples per step, BPR + EmbLoss, SOURCE phase.  A "step" is
one pass of the hot path over one batch: fused gather+score+loss forward, then re-gather+gradient+scatter-add
backward into the embedding-gradient tables (optimizer excluded, as in the metric's definition).

Printed JSON line (rank 0): the driver contract + `roofline` + `cpu_baseline` + `e2e` + `clocks` + `gpu_launches`.
  value     whole-job interactions/s with every input already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e       same metric through the public model API (EMCDR.calculate_loss + backward) fed from PINNED HOST batches:
            per step an H2D copy of the ids and a D2H read of the loss are inside the timed region
  roofline  HBM-bound; achieved = 1560 B/interaction x interactions per launch / launch duration (CUDA events)
  cpu_baseline  the oracle port of the reference step (oracle/cdr_oracle.py) on the host cores, bounded sample
`--impl reference` times that CPU port alone (the reference is pure Python on PyTorch and needs the un-vendored
recbole, so it cannot be installed on the GPU box; the oracle restates its arithmetic and cost structure: 6 gathers,
dense [N, D] gradients).  Inputs (0.9 GB of tables, random rows) are larger than the 126 MB L2; no extra L2 flush.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'recbole-cdr_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

BYTES_PER_INTERACTION_BPR_D64 = 3 * 8 + 3 * 4 * 64 + 3 * 4 * 64  # ids + gathered rows + scattered rows = 1560
# dram__bytes_read.sum + dram__bytes_write.sum of train_steps_staged_kernel<8,2,true> from the `ncu --set full` capture
# committed as profiles/r1_train_steps_staged_ncu.md (968.0 MB over a 60-step launch): 16.13 MB per 8192-interaction step
NCU_DRAM_BYTES_PER_STEP_B8192 = 968.0e6 / 60


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='xdr', choices=['xdr', 'reference'])
    ap.add_argument('--workload', default='emcdr_1m', choices=['emcdr_1m', 'emcdr_10m', 'emcdr_100k'])
    ap.add_argument('--batch', type=int, default=8192)
    ap.add_argument('--cpu-steps', type=int, default=3, help='timed steps of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-compare', action='store_true', help='skip the per-step-kernel comparison runs')
    ap.add_argument('--mode', default='persistent', choices=['persistent', 'per_step'])
    ap.add_argument('--shard-chunk', type=int, default=50, help='N>1: steps per persistent launch / peer-gather chunk')
    ap.add_argument('--stage-remote', action='store_true',
                    help='N>1: pull item rows with the peer-gather kernel one chunk ahead (slower in round 1, see profiles/)')
    ap.add_argument('--chunk', type=int, default=50, help='steps per launch on the end-to-end (host-fed) path')
    return ap.parse_args()


SCALES = {'emcdr_100k': 100_000, 'emcdr_1m': 1_000_000, 'emcdr_10m': 10_000_000}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's step, on the host cores
# ----------------------------------------------------------------------------------------------------------------------

def cpu_reference_step_rate(ds, batch, n_warm, n_timed, seed0=10_000):
    """EMCDR BPR SOURCE-phase step as the reference executes it (emcdr.py:121-130 + autograd dense grads):
    forward + backward to dense [N, D] embedding gradients.  Returns (interactions/s, seconds per step, cores)."""
    from oracle import cdr_oracle as O
    from recbole_cdr_b200.data import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(2022)
    ut = torch.nn.Parameter(O.xavier_normal_table(ds.num_total_user, 64, g))
    it = torch.nn.Parameter(O.xavier_normal_table(ds.num_total_item, 64, g))
    times = []
    for s in range(n_warm + n_timed):
        b = synthetic.make_batch(ds, 'source', batch, seed0 + s, 'cpu', pairwise=True)
        ut.grad = it.grad = None
        t0 = time.perf_counter()
        loss = O.emcdr_bpr_loss(ut, it, b['source_user_id'], b['source_item_id'], b['neg_source_item_id'], 0.01)
        loss.sum().backward()
        dt = time.perf_counter() - t0
        if s >= n_warm:
            times.append(dt)
    sec = sum(times) / len(times)
    return batch / sec, sec, cores


def run_reference(args):
    """`--impl reference`: rank 0 times the CPU port on the box's host cores; other ranks exit quietly."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    from recbole_cdr_b200.data import synthetic
    ds = synthetic.emcdr_scale(SCALES[args.workload])
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    rate, sec, cores = cpu_reference_step_rate(ds, args.batch, warm, steps)
    sample = (f'{steps} timed + {warm} warm-up steps of the same workload (B={args.batch}, tables {ds.num_total_user}x64 '
              f'and {ds.num_total_item}x64), forward+backward to dense grads, torch CPU {cores} threads')
    line = {
        'impl': 'reference', 'metric': 'interactions/sec (gather+map+score+scatter)', 'value': rate,
        'unit': 'interactions/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, ds),
        'cpu_baseline': {'value': rate, 'unit': 'interactions/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': 'interactions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(args, ds):
    return {'workload': f'EMCDR BPR SOURCE-phase step, synthetic {args.workload} (BASELINE configs[1] shape)',
            'users_total': ds.num_total_user, 'items_total': ds.num_total_item, 'dim': 64, 'batch_per_gpu': args.batch,
            'loss': 'BPR + 0.01*EmbLoss', 'l2': 'inputs larger than L2 (0.9 GB of tables, uniform random rows)',
            'parallelism': (f'tables row-sharded over {args.gpus} GPUs (r mod G) on peer memory, batch data-parallel and '
                            'routed by user owner; gathers/scatter-adds cross NVLink inside the kernel, no collective on '
                            'the data path') if args.gpus > 1 else 'single GPU'}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------

def run_xdr(args):
    import torch.distributed as dist
    from recbole_cdr_b200 import _lib, ops
    from recbole_cdr_b200.data import synthetic
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl xdr needs a CUDA device (the hot path has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    ds = synthetic.emcdr_scale(SCALES[args.workload])
    B, K, W, D = args.batch, args.steps, args.warmup, 64
    cfg = {'source_domain': {'NEG_PREFIX': 'neg_'}, 'target_domain': {'NEG_PREFIX': 'neg_'}, 'device': dev,
           'latent_factor_model': 'BPR', 'source_embedding_size': D, 'target_embedding_size': D, 'reg_weight': 0.01,
           'mapping_function': 'non_linear', 'mlp_hidden_size': [128]}
    sharded = world > 1
    if not sharded:
        torch.manual_seed(2022)
        with torch.device(dev):
            model = EMCDR(cfg, ds)  # random-init weights of the named architecture, created directly in HBM
        model.set_phase('SOURCE')
        ut, it = model.source_user_embedding.weight, model.source_item_embedding.weight
        gu, gi = torch.zeros_like(ut), torch.zeros_like(it)
    else:
        # row-sharded tables (block-cyclic, r mod G) mapped over CUDA IPC; every rank holds 1/G of each table and of
        # each gradient table; xavier-normal random init of the named shapes, created directly in HBM
        from recbole_cdr_b200 import shard
        torch.manual_seed(2022 + rank)
        def mk(n, fill):
            rows = shard.shard_rows(n, world)
            loc = (torch.randn(rows, D, device=dev) * (2.0 / (n + D)) ** 0.5) if fill else torch.zeros(rows, D, device=dev)
            return shard.RowShardedTable(n, D, rank, world, dev, loc).connect()
        s_ut, s_it = mk(ds.num_total_user, True), mk(ds.num_total_item, True)
        s_gu, s_gi = mk(ds.num_total_user, False), mk(ds.num_total_item, False)
        sh_runner = shard.ShardedStepRunner(s_ut, s_it, s_gu, s_gi, reg_weight=0.01, chunk=args.shard_chunk,
                                            stage_remote=args.stage_remote)
        dist.barrier()

    # K + W distinct seeded batches (seed = 1 + step, offset per rank), resident in HBM and mirrored in pinned host memory
    def batch_ids(step):
        b = synthetic.make_batch(ds, 'source', B, 1 + step + 100_003 * rank, 'cpu', pairwise=True)
        u = b['source_user_id']
        if sharded:  # the loader routes an interaction to the rank that owns its user row: user % G == rank
            u = u - ((u - rank) % world)
            u = torch.where(u < 1, u + world, u)
        return torch.stack([u, b['source_item_id'], b['neg_source_item_id']])

    host = torch.stack([batch_ids(s) for s in range(K + W)]).pin_memory()  # [K+W, 3, B] int64
    ids = host.to(dev)
    scores = torch.empty(2, B, device=dev)
    out8 = torch.empty(K + W, 8, device=dev)
    ws = _lib.workspace(dev)
    stream = _lib.cur_stream()

    def step(s):
        u, ip, ineg = ids[s, 0], ids[s, 1], ids[s, 2]
        _lib.call('xdr_bpr_fwd', ut.data_ptr(), it.data_ptr(), ut.shape[0], it.shape[0], D, u.data_ptr(), ip.data_ptr(),
                  ineg.data_ptr(), B, 1e-10, 0.01, scores[0].data_ptr(), scores[1].data_ptr(), out8[s].data_ptr(),
                  ws.data_ptr(), None, _lib.cur_stream())
        _lib.call('xdr_bpr_bwd', ut.data_ptr(), it.data_ptr(), ut.shape[0], it.shape[0], D, u.data_ptr(), ip.data_ptr(),
                  ineg.data_ptr(), B, 1e-10, 0.01, scores[0].data_ptr(), scores[1].data_ptr(), out8[s].data_ptr(), None,
                  1.0, gu.data_ptr(), gi.data_ptr(), _lib.cur_stream())

    def persistent(lo, hi):
        """steps [lo, hi) as ONE persistent launch (xdr_train_steps): fwd + bwd + scatter-add per batch"""
        if sharded:
            sh_runner.run(ids[lo:hi], out8=out8[lo:hi])
        else:
            ops.train_steps(ut.data, it.data, ids[lo:hi, 0], ids[lo:hi, 1], ids[lo:hi, 2], reg_weight=0.01, user_dst=gu,
                            item_dst=gi, out8=out8[lo:hi])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_persistent = sharded or (args.mode == 'persistent' and ops.train_steps_supported(B, D, True, dev))
    # ---- device-resident throughput -----------------------------------------------------------------------------
    if use_persistent:
        persistent(0, W)
        if sharded:
            sh_runner.launches = 0
    else:
        for s in range(W):
            step(s)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if use_persistent:
        persistent(W, W + K)
        launches = 1 if not sharded else sh_runner.launches
    else:
        for s in range(W, W + K):
            step(s)
        launches = 2 * K
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    loss_mean = float(out8[W:, 0].mean().item())

    extra = {}
    # ---- same launch with the scatter-add aimed at the weight tables (scale = -lr): the SGD update fused into the step
    if use_persistent and not sharded:
        ops.train_steps(ut.data, it.data, ids[:W, 0], ids[:W, 1], ids[:W, 2], reg_weight=0.01, user_dst=ut.data,
                        item_dst=it.data, scale=-1e-3, out8=out8[:W])
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        ops.train_steps(ut.data, it.data, ids[W:, 0], ids[W:, 1], ids[W:, 2], reg_weight=0.01, user_dst=ut.data,
                        item_dst=it.data, scale=-1e-3, out8=out8[W:])
        f1.record()
        barrier()
        fms = torch.tensor([f0.elapsed_time(f1)], device=dev)
        if world > 1:
            dist.all_reduce(fms, op=dist.ReduceOp.MAX)
        fused_rate = world * B * K / (fms.item() * 1e-3)
        extra['fused_sgd'] = {
            'note': 'same persistent launch, scatter-add of -lr*grad straight into the embedding tables (row-sparse SGD '
                    'step included; rows are L2-resident for the atomics so DRAM traffic ~= algorithmic bytes)',
            'value': fused_rate, 'ms_per_step': fms.item() / K,
            'roofline_frac': BYTES_PER_INTERACTION_BPR_D64 * fused_rate / world / 1e9 / measured_peaks()[0]}
    # ---- for comparison: the per-step kernel pair replayed from a CUDA graph (no host launch cost) -----------------
    if use_persistent and not sharded and rank == 0 and not args.no_compare:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for s in range(3):
                step(s)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                for s in range(W, W + K):
                    step(s)
        torch.cuda.synchronize()
        graph.replay()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        graph.replay()
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1)
        extra['per_step_kernels_cuda_graph'] = {'ms_per_step': gms / K, 'value': B * K / (gms * 1e-3),
                                                'launches': 2 * K}
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for s in range(W, W + K):
            step(s)
        p1.record()
        torch.cuda.synchronize()
        pms = p0.elapsed_time(p1)
        extra['per_step_kernels_eager'] = {'ms_per_step': pms / K, 'value': B * K / (pms * 1e-3), 'launches': 2 * K}

    # ---- end-to-end through the public trainer API, host batches ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        from recbole_cdr_b200.trainer import FusedStepRunner
        chunk = min(args.chunk, K)
        n_chunks = K // chunk
        if sharded:
            def launch(idb, _label, o8):
                sh_runner.run(idb, out8=o8)
            runner = FusedStepRunner({'pairwise': True}, launch=launch, device=dev)
        else:
            runner = FusedStepRunner(model.fused_step_spec(), lr=None, grad_tables=(gu, gi))
        blocks = [host[W + c * chunk: W + (c + 1) * chunk] for c in range(n_chunks)]  # pinned [chunk, 3, B] views
        for c in range(min(3, n_chunks)):
            runner.run(blocks[c])
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        losses = [runner.run(b) for b in blocks]  # per chunk: H2D ids -> one persistent launch -> D2H losses
        a1.record()
        barrier()
        ms_api = a0.elapsed_time(a1)
        loss_e2e = float(torch.stack([l.mean() for l in losses]).mean())
        e2e_ms = torch.tensor([ms_api], device=dev)
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e = {'value': world * B * chunk * n_chunks / (e2e_ms.item() * 1e-3), 'unit': 'interactions/s',
               'h2d_bytes_per_step': 3 * 8 * B, 'd2h_bytes_per_step': 4, 'steps': chunk * n_chunks,
               'steps_per_launch': chunk, 'loss_mean': loss_e2e,
               'api': 'trainer.FusedStepRunner.run(pinned [chunk,3,B] int64 ids): H2D copy -> xdr_train_steps -> D2H losses'}
    clk = clocks.stop() if rank == 0 else None

    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()

    if rank == 0:
        peak, peak_src = measured_peaks()
        value = world * B * K / (ms * 1e-3)
        # persistent mode: ONE launch processes K*B interactions in `ms`; per-step mode: a step = fwd + bwd launch pair
        achieved = BYTES_PER_INTERACTION_BPR_D64 * B * K / (ms * 1e-3) / 1e9
        line = {
            'metric': 'interactions/sec (gather+map+score+scatter)', 'value': value, 'unit': 'interactions/s',
            'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, ds),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': (NCU_DRAM_BYTES_PER_STEP_B8192 * K if (use_persistent and B == 8192 and world == 1) else None),
                         'traffic_note': 'bytes per launch = ncu dram read+write per step (profiles/r1_train_steps_staged_ncu.md) x K',
                         'peak_source': peak_src,
                         'kernel': ('train_steps_staged_kernel<8,2,true>: one persistent launch over all K timed steps'
                                    if use_persistent else 'score_fwd_kernel<2,true> + score_bwd_kernel<2,true> per step'),
                         'units_per_launch': B * K if use_persistent else B,
                         'bytes_per_interaction': BYTES_PER_INTERACTION_BPR_D64},
            'e2e': e2e, 'gpu_launches': launches, 'clocks': clk, 'loss_mean': loss_mean,
        }
        line.update(extra)
        if not args.no_cpu_baseline and world == 1:
            rate, sec, cores = cpu_reference_step_rate(ds, B, 1, args.cpu_steps)
            line['cpu_baseline'] = {
                'value': rate, 'unit': 'interactions/s', 'cores': cores, 'kind': 'port',
                'sample': f'{args.cpu_steps} timed + 1 warm-up steps of the same workload on the host (oracle port of '
                          f'emcdr.py:121-130 + autograd dense grads), {sec * 1e3:.0f} ms/step'}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_xdr(args)


if __name__ == '__main__':
    main()
