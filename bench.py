#!/usr/bin/env python
"""bench.py -- interactions/s through the gather -> map -> score -> scatter hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl xdr|reference] [--workload ...]

Workloads (config.workload):
  emcdr_1m   BASELINE.json configs[1] -- EMCDR, synthetic 1M x 1M users/items per domain, dim 64, batch 8192 (user, item+,
             item-) triples per step, BPR + 0.01 EmbLoss, SOURCE phase.  Default at N = 1.
  emcdr_10m  configs[4] -- the same step on 10M x 10M tables, batch 8192 per GPU.  Default at N > 1 (tables row-sharded over
             the GPUs); at N = 1 it is measured as the `emcdr_10m` sub-object of the line (the denominator of the 8-GPU claim).
  (the default N = 1 line also carries `model_steps`: the emcdr_map and conet_5m lines below, each measured by this script in a
   child process after the main numbers are final)
  emcdr_map  the OVERLAP-phase mapping step (gather -> MLP 64-128-64 -> MSE -> backward -> scatter), b = 8192.
A "step" is one pass of the hot path over one batch: gather + score + loss forward, row gradients, scatter-add into the
embedding-gradient tables (optimizer excluded, as in the metric's definition).

Timing.  The K steps of the timed region are ONE persistent launch (xdr_train_steps).  Every sample is taken behind a
GPU-side gate: a spin kernel (torch.cuda._sleep) is queued first, the start event, the launch and the stop event are enqueued
while it spins, so host enqueue time is not inside the event pair.  The K-step launch is repeated R times (--repeats) on
different batches; `ms_per_step` / `value` are the MEDIAN sample (min / max in `timing`); ranks are synchronised with a
barrier + cuda synchronize on both sides of every sample and every sample is the max over ranks.

Printed JSON line (rank 0): the driver contract + `roofline` + `cpu_baseline` + `e2e` + `clocks` + `gpu_launches`.
  value     whole-job interactions/s with every input already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e       the same metric through the public trainer API (trainer.FusedStepRunner.run) fed from PINNED HOST id blocks:
            per chunk an H2D copy of the ids, one launch and a D2H read of the per-step losses are inside the timed region
  roofline  HBM-bound; achieved = 1560 B/interaction x interactions per launch / launch duration (CUDA events)
  cpu_baseline / --impl reference: the UNMODIFIED reference EMCDR class (staged by build() into the git-ignored oracle/_ref
            next to oracle/recbole_shim, the stub of the un-vendored recbole) on the host cores: (a) calculate_loss forward,
            (b) forward + backward, (c) + dense Adam step -- (b) is the line's value (the metric excludes the optimizer).
            Falls back to the oracle port (kind "port") only if oracle/_ref is absent.
Inputs (0.9 GB of tables, random rows) are larger than the 126 MB L2; no extra L2 flush.
"""
import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'recbole-cdr_b200')

import torch  # noqa: E402

D = 64
BYTES_PER_INTERACTION_BPR_D64 = 3 * 8 + 3 * 4 * D + 3 * 4 * D   # ids + gathered rows + scattered rows = 1560
BYTES_PER_ROW_MAP_D64 = 8 + 2 * 4 * D + 2 * 4 * D               # 1 id + 2 rows gathered + 2 rows scattered = 1032
SCALES = {'emcdr_100k': 100_000, 'emcdr_1m': 1_000_000, 'emcdr_10m': 10_000_000}
GATE_CYCLES = 600_000    # ~0.3 ms of GPU-side spinning in front of every timed sample


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='xdr', choices=['xdr', 'reference'])
    ap.add_argument('--workload', default=None, choices=['emcdr_1m', 'emcdr_10m', 'emcdr_100k', 'emcdr_map', 'conet_5m'])
    ap.add_argument('--batch', type=int, default=8192)
    ap.add_argument('--repeats', type=int, default=11, help='timed K-step launches (median reported)')
    ap.add_argument('--cpu-steps', type=int, default=3, help='timed steps of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip variants / fused SGD / per-step comparison / 10M sub-run')
    ap.add_argument('--grad-mode', default='accumulate', choices=['fresh', 'accumulate'],
                    help='accumulate (default): scatter-add into the dense gradient tables; fresh: lazily zeroed gradient tables -- '
                         'every launch produces the gradient of its K batches, first touch of a row stores (touch map)')
    ap.add_argument('--coop', type=int, default=1, help='1: cudaLaunchCooperativeKernel, 0: plain launch')
    ap.add_argument('--dense-engine', type=int, default=-1,
                    help='model-step workloads: 0 fp32 FMA dense layers, 1 tcgen05, 2 tcgen05 where it measured faster (wide layers, '
                         'forward / input gradient), -1 the library default')
    ap.add_argument('--map-engine', default='', help="emcdr_map: '' (the model's default: tcgen05), 'tc5', 'tc' (mma.sync), 'fma', 'composed'")
    ap.add_argument('--shard-chunk', type=int, default=50, help='N>1: steps per persistent launch')
    ap.add_argument('--chunk', type=int, default=0, help='steps per launch on the end-to-end (host-fed) path (0: K/2, 50 from K = 100)')
    return ap.parse_args()


def add_paths():
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)


def load_synthetic():
    """recbole_cdr_b200.data.{idspace,synthetic} WITHOUT importing the package (whose __init__ dlopens libxdr.so): the
    reference arm must not load any of this repository's native code."""
    name = '_xdr_data_only'
    if name + '.synthetic' in sys.modules:
        return sys.modules[name + '.synthetic']
    ddir = os.path.join(PKG, 'recbole_cdr_b200', 'data')
    pkg = types.ModuleType(name)
    pkg.__path__ = [ddir]
    sys.modules[name] = pkg
    for sub in ('idspace', 'synthetic'):
        spec = importlib.util.spec_from_file_location(f'{name}.{sub}', os.path.join(ddir, sub + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f'{name}.{sub}'] = mod
        spec.loader.exec_module(mod)
    return sys.modules[name + '.synthetic']


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
# CPU arm: the reference's own EMCDR class (oracle/_ref) on the host cores; the oracle port only as a fallback
# ----------------------------------------------------------------------------------------------------------------------

def _import_reference_emcdr():
    """The unmodified reference class from oracle/_ref (staged by __graft_entry__.build(), git-ignored) over the recbole stub."""
    ref = os.path.join(ROOT, 'oracle', '_ref')
    if not os.path.exists(os.path.join(ref, 'recbole_cdr', 'model', 'cross_domain_recommender', 'emcdr.py')):
        return None
    for p in (os.path.join(ROOT, 'oracle', 'recbole_shim'), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    if not hasattr(np, 'NINF'):
        np.NINF = -np.inf   # NumPy-2 compatibility shim for the reference (dtcdr.py:55-59)
    from recbole_cdr.model.cross_domain_recommender.emcdr import EMCDR
    return EMCDR


def cpu_reference_rates(scale, batch, n_warm, n_timed, with_adam=True, seed0=10_000):
    """The reference step on the host cores.  Returns dict(kind, cores, fwd, fwd_bwd, fwd_bwd_adam) with interactions/s and
    ms per step for (a) EMCDR.calculate_loss forward, (b) + backward (dense [N, D] embedding gradients), (c) + torch.optim.Adam
    step (the reference's default learner, overall.yaml:20-21) -- BASELINE.md section 4."""
    synthetic = load_synthetic()
    ds = synthetic.emcdr_scale(scale)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    EMCDR = _import_reference_emcdr()
    out = {'cores': cores}
    if EMCDR is not None:
        cfg = {'source_domain': {'NEG_PREFIX': 'neg_'}, 'target_domain': {'NEG_PREFIX': 'neg_'}, 'device': 'cpu',
               'latent_factor_model': 'BPR', 'source_embedding_size': D, 'target_embedding_size': D, 'reg_weight': 0.01,
               'mapping_function': 'non_linear', 'mlp_hidden_size': [128], 'overlap_batch_size': 100}
        torch.manual_seed(2022)
        model = EMCDR(cfg, ds)
        model.set_phase('SOURCE')
        out['kind'] = 'reference'
        loss_fn = model.calculate_loss
        params = [p for p in model.parameters()]
    else:
        add_paths()
        from oracle import cdr_oracle as O
        g = torch.Generator().manual_seed(2022)
        ut = torch.nn.Parameter(O.xavier_normal_table(ds.num_total_user, D, g))
        it = torch.nn.Parameter(O.xavier_normal_table(ds.num_total_item, D, g))
        out['kind'] = 'port'
        params = [ut, it]

        def loss_fn(b):
            return O.emcdr_bpr_loss(ut, it, b['source_user_id'], b['source_item_id'], b['neg_source_item_id'], 0.01)

    def batch_of(s):
        return synthetic.make_batch(ds, 'source', batch, seed0 + s, 'cpu', pairwise=True)

    def run(kind, n_w, n_t):
        opt = torch.optim.Adam(params, lr=1e-3) if kind == 'adam' else None
        ts = []
        for s in range(n_w + n_t):
            b = batch_of(s)
            t0 = time.perf_counter()
            if kind == 'fwd':
                with torch.no_grad():
                    loss_fn(b)
            else:
                if opt is not None:
                    opt.zero_grad()
                else:
                    for p in params:
                        p.grad = None
                loss = loss_fn(b)
                loss.sum().backward()
                if opt is not None:
                    opt.step()
            dt = time.perf_counter() - t0
            if s >= n_w:
                ts.append(dt)
        sec = sum(ts) / len(ts)
        return {'value': batch / sec, 'ms_per_step': sec * 1e3}

    out['fwd'] = run('fwd', max(1, n_warm), max(n_timed, 5))
    out['fwd_bwd'] = run('bwd', n_warm, n_timed)
    if with_adam:
        out['fwd_bwd_adam'] = run('adam', 1, max(1, min(n_timed, 2)))
    return out


def cpu_baseline_object(r, batch, scale, n_timed):
    ds_rows = {100_000: (150_001, 200_001), 1_000_000: (1_500_001, 2_000_001), 10_000_000: (15_000_001, 20_000_001)}[scale]
    what = ('the unmodified reference EMCDR class (oracle/_ref/recbole_cdr over oracle/recbole_shim)' if r['kind'] == 'reference'
            else 'oracle port of emcdr.py:121-130 (oracle/_ref not staged)')
    o = {'value': r['fwd_bwd']['value'], 'unit': 'interactions/s', 'cores': r['cores'], 'kind': r['kind'],
         'sample': (f'{n_timed} timed + 1 warm-up steps of the same workload (B={batch}, tables {ds_rows[0]}x64 and '
                    f'{ds_rows[1]}x64, 4 tables allocated as the reference does), {what}, torch CPU {r["cores"]} threads; '
                    f'value = (b) forward+backward to dense grads, {r["fwd_bwd"]["ms_per_step"]:.0f} ms/step'),
         'forward': r['fwd'], 'forward_backward': r['fwd_bwd']}
    if 'fwd_bwd_adam' in r:
        o['forward_backward_adam'] = r['fwd_bwd_adam']
    return o


def run_reference(args):
    """`--impl reference`: rank 0 times the reference's CPU implementation on the box's host cores; other ranks exit quietly."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    workload = args.workload or ('emcdr_1m' if args.gpus == 1 else 'emcdr_10m')
    if workload in ('emcdr_map', 'conet_5m'):
        workload = 'emcdr_1m'
    scale = SCALES[workload]
    steps = max(1, min(args.steps, 3 if scale > 1_000_000 else 5))
    r = cpu_reference_rates(scale, args.batch, 1, steps, with_adam=scale <= 1_000_000)
    synthetic = load_synthetic()
    ds = synthetic.emcdr_scale(scale)
    cb = cpu_baseline_object(r, args.batch, scale, steps)
    rate = r['fwd_bwd']['value']
    line = {
        'impl': 'reference', 'metric': 'interactions/sec (gather+map+score+scatter)', 'value': rate,
        'unit': 'interactions/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': 1, 'ms_per_step': r['fwd_bwd']['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(workload, args.batch, args.gpus, ds),
        'cpu_baseline': cb,
        'e2e': {'value': rate, 'unit': 'interactions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(workload, batch, gpus, ds, extra=None):
    which = {'emcdr_1m': 'BASELINE configs[1] shape', 'emcdr_10m': 'BASELINE configs[4] shape, batch 8192 per GPU',
             'emcdr_100k': 'reduced smoke shape'}.get(workload, workload)
    gb = (ds.num_total_user + ds.num_total_item) * D * 4 / 1e9
    cfg = {'workload': f'EMCDR BPR SOURCE-phase step, synthetic {workload} ({which})',
           'users_total': ds.num_total_user, 'items_total': ds.num_total_item, 'dim': D, 'batch_per_gpu': batch,
           'loss': 'BPR + 0.01*EmbLoss', 'l2': f'inputs larger than L2 ({gb:.1f} GB of tables, uniform random rows)',
           'parallelism': (f'tables row-sharded over {gpus} GPUs (r mod G) on peer memory, batch data-parallel and '
                           'routed by user owner; gathers/scatter-adds cross NVLink inside the kernel, no collective on '
                           'the data path') if gpus > 1 else 'single GPU'}
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------

class Timer:
    """Gate-timed samples of `fn(r)` (r = repeat index): median / min / max in ms, max over ranks per sample."""

    def __init__(self, dev, world):
        self.dev, self.world = dev, world

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def samples(self, fn, repeats):
        out = []
        for r in range(repeats):
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(GATE_CYCLES)   # GPU-side gate: everything below is enqueued while the device spins
            e0.record()
            fn(r)
            e1.record()
            self.barrier()
            out.append(e0.elapsed_time(e1))
        t = torch.tensor(out, device=self.dev, dtype=torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    @staticmethod
    def summary(ms):
        return {'repeats': len(ms), 'median_ms': statistics.median(ms), 'min_ms': min(ms), 'max_ms': max(ms),
                'gate': 'spin kernel queued before the start event (host enqueue time excluded)'}


class RecWorkload:
    """EMCDR SOURCE-phase BPR step on one GPU: tables + gradient tables in HBM, (W + R*K) seeded batches resident."""

    def __init__(self, scale, B, K, W, R, dev, std=None, zipf=None, grad_mode='fresh', seed_off=0, hot_rows=False):
        from recbole_cdr_b200.data import synthetic
        self.ds = synthetic.emcdr_scale(scale)
        self.B, self.K, self.W, self.R, self.dev, self.grad_mode = B, K, W, R, dev, grad_mode
        g = torch.Generator(device=dev).manual_seed(2022)
        nu, ni = self.ds.num_total_user, self.ds.num_total_item
        su = std if std is not None else (2.0 / (nu + D)) ** 0.5     # xavier_normal_ of an [N, D] table
        si = std if std is not None else (2.0 / (ni + D)) ** 0.5
        self.ut = torch.randn(nu, D, device=dev, generator=g) * su
        self.it = torch.randn(ni, D, device=dev, generator=g) * si
        self.gu, self.gi = torch.zeros_like(self.ut), torch.zeros_like(self.it)
        n = W + R * K
        ids = torch.empty(n, 3, B, dtype=torch.int64, device=dev)
        for s in range(n):
            b = synthetic.make_batch(self.ds, 'source', B, 1 + s + seed_off, dev, pairwise=True, zipf_items=zipf)
            ids[s, 0], ids[s, 1], ids[s, 2] = b['source_user_id'], b['source_item_id'], b['neg_source_item_id']
        self.ids = ids
        self.out8 = torch.empty(n, 8, device=dev)
        self.touch = None
        # popular rows (a dataset statistic, computed once): pre-aggregated per CTA in shared memory (xdr_steps_set_hot_rows)
        self.hot = None
        if hot_rows:
            from recbole_cdr_b200 import ops
            self.hot = (ops.hot_rows_from_ids(ids[:, 0], nu, 16), ops.hot_rows_from_ids(ids[:, 1:], ni, 48))

    def touch_map(self):
        from recbole_cdr_b200 import ops
        if self.touch is None:
            self.touch = ops.TouchMap(self.ut.shape[0], self.it.shape[0], self.dev)
        return self.touch

    def launch(self, lo, hi, dst=None, scale=1.0):
        from recbole_cdr_b200 import ops
        ids = self.ids
        kw = {}
        if dst is None:
            dst = (self.gu, self.gi)
            if self.grad_mode == 'fresh':   # lazily zeroed gradient tables: this launch's gradient, no RMW of gradient lines
                kw = dict(touch=self.touch_map(), fresh=True)
        if self.hot is not None:
            ops.set_steps_hot_rows(*self.hot)
        try:
            ops.train_steps(self.ut, self.it, ids[lo:hi, 0], ids[lo:hi, 1], ids[lo:hi, 2], reg_weight=0.01, user_dst=dst[0],
                            item_dst=dst[1], scale=scale, out8=self.out8[lo:hi], **kw)
        finally:
            if self.hot is not None:
                ops.set_steps_hot_rows(None, None)

    def timed(self, timer, dst=None, scale=1.0):
        K, W = self.K, self.W
        self.launch(0, W, dst, scale)
        ms = timer.samples(lambda r: self.launch(W + r * K, W + (r + 1) * K, dst, scale), self.R)
        return ms, float(self.out8[W:, 0].mean().item())


def sequential_row_sparse(wl, n_steps, B):
    """n_steps strictly sequential (gradient -> optimizer) steps, eager launches, CUDA events around the loop."""
    from recbole_cdr_b200 import ops
    from recbole_cdr_b200.trainer.row_optim import RowSparseOptimizer
    out = {}
    ut = torch.nn.Parameter(wl.ut.clone())
    it = torch.nn.Parameter(wl.it.clone())
    ut.grad, it.grad = torch.zeros_like(ut), torch.zeros_like(it)
    ids = wl.ids
    for kind in ('adagrad', 'lazy_adam'):
        opt = RowSparseOptimizer(kind, lr=1e-3)

        def step(k):
            ops.train_steps(ut.data, it.data, ids[k:k + 1, 0], ids[k:k + 1, 1], ids[k:k + 1, 2], reg_weight=0.01,
                            user_dst=ut.grad, item_dst=it.grad)
            opt.step([(ut, ids[k, 0]), (it, ids[k, 1:])])
        for k in range(3):
            step(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n_steps):
            step(3 + k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_steps
        out[kind] = {'ms_per_step': ms, 'value': B / (ms * 1e-3)}
    # dense reference semantics on the GPU: zero_grad + (gradient as above) + torch.optim.Adam over the whole tables
    dopt = torch.optim.Adam([ut, it], lr=1e-3)

    def dense_step(k):
        ut.grad.zero_()
        it.grad.zero_()
        ops.train_steps(ut.data, it.data, ids[k:k + 1, 0], ids[k:k + 1, 1], ids[k:k + 1, 2], reg_weight=0.01,
                        user_dst=ut.grad, item_dst=it.grad)
        dopt.step()
    for k in range(2):
        dense_step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(5):
        dense_step(2 + k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out['dense_torch_adam'] = {'ms_per_step': ms, 'value': B / (ms * 1e-3)}
    out['note'] = ('sequential semantics (weights updated after EVERY batch): persistent launch of one step + row-sparse optimizer '
                   'kernels over the batch ids, eager Python launches (host-bound); dense_torch_adam = the same gradient launch + '
                   'zero_grad + torch.optim.Adam over the full tables')
    del ut, it, dopt
    torch.cuda.empty_cache()
    return out


def rate_fields(ms_list, K, B, world, peak):
    med = statistics.median(ms_list)
    value = world * B * K / (med * 1e-3)
    ach = BYTES_PER_INTERACTION_BPR_D64 * B * K / (med * 1e-3) / 1e9
    return {'value': value, 'ms_per_step': med / K, 'roofline_frac': ach / peak, 'min_ms_per_step': min(ms_list) / K,
            'max_ms_per_step': max(ms_list) / K}


def sharded_parity_check(rank, world, dev):
    """Small-size equality of the row-sharded step with the single-GPU kernel, run inside the bench so that every multi-GPU
    line carries parity evidence: per-step losses of this rank's batches and this rank's rows of the accumulated gradient
    tables against one single-GPU launch over the union of all ranks' batches."""
    import torch.distributed as dist
    from recbole_cdr_b200 import ops, shard
    nu, ni, B, K = 4099, 6151, 512, 3
    g = torch.Generator().manual_seed(99)
    ut, it = (torch.randn(nu, D, generator=g) * 0.1).to(dev), (torch.randn(ni, D, generator=g) * 0.1).to(dev)
    ids = torch.stack([torch.stack([torch.randint(1, nu, (K, B), generator=g), torch.randint(1, ni, (K, B), generator=g),
                                    torch.randint(1, ni, (K, B), generator=g)], 1) for _ in range(world)]).to(dev)  # [G,K,3,B]
    tabs = [shard.RowShardedTable.from_full(t, rank, world, dev).connect() for t in (ut, it)]
    dsts = [shard.RowShardedTable(n, D, rank, world, dev).connect() for n in (nu, ni)]
    dist.barrier()
    mine = ids[rank].contiguous()
    out8 = shard.train_steps_sharded(tabs[0], tabs[1], dsts[0], dsts[1], mine[:, 0], mine[:, 1], mine[:, 2], reg_weight=0.01)
    torch.cuda.synchronize()
    dist.barrier()
    gu, gi = torch.zeros_like(ut), torch.zeros_like(it)
    losses = None
    for r in range(world):
        b = ids[r].contiguous()
        o, _, _ = ops.train_steps(ut, it, b[:, 0], b[:, 1], b[:, 2], reg_weight=0.01, user_dst=gu, item_dst=gi)
        if r == rank:
            losses = o[:, 0].clone()
    torch.cuda.synchronize()
    ok_loss = bool(torch.equal(out8[:, 0], losses))
    errs = []
    for full, sh in ((gu, dsts[0]), (gi, dsts[1])):
        want = full[rank::world]
        got = sh.local[:want.shape[0]]
        errs.append(float((got - want).abs().max() / want.abs().max().clamp_min(1e-30)))
    flag = torch.tensor([1.0 if (ok_loss and max(errs) < 1e-5) else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    for t in tabs + dsts:
        t.close()
    res = {'losses_bit_equal': ok_loss, 'grad_max_rel_err': max(errs), 'all_ranks_ok': bool(flag.item() == 1.0),
           'what': f'{K} steps x B={B} per rank on {nu}x{D} / {ni}x{D} tables: sharded launch vs single-GPU launches over the '
                   'union of the ranks\' batches (losses bit-equal, gradient rows within 1e-5 of the table max)'}
    if not res['all_ranks_ok']:
        raise RuntimeError(f'sharded parity check failed on rank {rank}: {res}')
    return res


def run_xdr(args):
    add_paths()
    import torch.distributed as dist
    from recbole_cdr_b200 import _lib, ops
    from recbole_cdr_b200.data import synthetic

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl xdr needs a CUDA device (the hot path has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    _lib._lib.xdr_set_coop_launch(int(args.coop))
    workload = args.workload or ('emcdr_1m' if world == 1 else 'emcdr_10m')
    if workload in MODEL_WORKLOADS:
        return run_model_step(args, dev, workload)
    scale = SCALES[workload]
    B, K, W, R = args.batch, args.steps, max(3, args.warmup), max(1, args.repeats)
    peak, peak_src = measured_peaks()
    timer = Timer(dev, world)
    sharded = world > 1
    extra = {}

    if not sharded:
        wl = RecWorkload(scale, B, K, W, R, dev, grad_mode=args.grad_mode)
        ds = wl.ds
        assert ops.train_steps_supported(B, D, True, dev)
        clocks = ClockSampler(local)
        clocks.start()
        ms_list, loss_mean = wl.timed(timer)
        launches = 1
        kernel_name = 'train_steps_staged_kernel<8,2,true>: one persistent launch over all K timed steps'
    else:
        # row-sharded tables (block-cyclic, r mod G) mapped over CUDA IPC; every rank holds 1/G of each table and of
        # each gradient table; xavier-normal random init of the named shapes, created directly in HBM
        from recbole_cdr_b200 import shard
        extra['sharded_parity'] = sharded_parity_check(rank, world, dev)
        ds = synthetic.emcdr_scale(scale)
        torch.manual_seed(2022 + rank)

        def mk(n, fill):
            rows = shard.shard_rows(n, world)
            loc = (torch.randn(rows, D, device=dev) * (2.0 / (n + D)) ** 0.5) if fill else torch.zeros(rows, D, device=dev)
            return shard.RowShardedTable(n, D, rank, world, dev, loc).connect()
        s_ut, s_it = mk(ds.num_total_user, True), mk(ds.num_total_item, True)
        s_gu, s_gi = mk(ds.num_total_user, False), mk(ds.num_total_item, False)
        sh_runner = shard.ShardedStepRunner(s_ut, s_it, s_gu, s_gi, reg_weight=0.01, chunk=max(args.shard_chunk, K))
        dist.barrier()
        n = W + R * K
        ids = torch.empty(n, 3, B, dtype=torch.int64, device=dev)
        for s in range(n):
            # the loader partitions the interaction stream by the owner of the user row (user % G == rank): every rank draws
            # 2B candidates of the global stream for this step and keeps the first B it owns
            kept, seed = [], 1 + s
            while sum(k.shape[1] for k in kept) < B:
                b = synthetic.make_batch(ds, 'source', 2 * B * world, seed, dev, pairwise=True)
                c = torch.stack([b['source_user_id'], b['source_item_id'], b['neg_source_item_id']])
                kept.append(c[:, (c[0] % world) == rank])
                seed += 1_000_003
            ids[s] = torch.cat(kept, 1)[:, :B]
        out8 = torch.empty(n, 8, device=dev)
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        sh_runner.run(ids[:W], out8=out8[:W])
        sh_runner.launches = 0
        ms_list = timer.samples(lambda r: sh_runner.run(ids[W + r * K: W + (r + 1) * K], out8=out8[W + r * K: W + (r + 1) * K]), R)
        launches = sh_runner.launches // R
        loss_mean = float(out8[W:, 0].mean().item())
        kernel_name = 'train_steps_staged_kernel<8,2,true> over peer-mapped shards: one persistent launch over all K timed steps'

    # ---- explanatory extras (single GPU) ---------------------------------------------------------------------------
    if not sharded and not args.no_extras:
        Re = max(3, R // 2)
        # the other gradient-destination mode of the same launch
        other = 'accumulate' if args.grad_mode == 'fresh' else 'fresh'
        wl.grad_mode, wl.R = other, Re
        ms_o, _ = wl.timed(timer)
        extra['grad_mode_' + other] = dict(rate_fields(ms_o, K, B, 1, peak), note=(
            'scatter-add into whatever the dense gradient tables hold (every RED into a DRAM-resident line is a '
            'read-modify-write)' if other == 'accumulate' else 'touch map: first touch of a row zero-fills it'))
        wl.grad_mode, wl.R = args.grad_mode, R
        # scatter-add aimed at the weight tables (scale = -lr): the SGD update fused into the step
        wl.R = Re
        ms_f, _ = wl.timed(timer, dst=(wl.ut, wl.it), scale=-1e-3)
        extra['fused_sgd'] = dict(rate_fields(ms_f, K, B, 1, peak), note=(
            'same persistent launch, scatter-add of -lr*grad straight into the embedding tables (row-sparse asynchronous SGD '
            'step included; rows are L2-resident for the atomics so DRAM traffic ~= algorithmic bytes)'))
        wl.R = R
        # per-step kernel pair replayed from a CUDA graph (sequential-semantics path, no host launch cost)
        extra.update(per_step_comparison(wl, K, W, B, peak))
        # strictly sequential training steps WITH an optimizer (verdict r1, item 7): per batch one persistent launch (K = 1)
        # into the gradient tables + the row-sparse optimizer kernel over the batch's ids (which re-zeroes the rows it
        # consumed) -- against dense torch.optim.Adam over the same tables, which is what the reference's trainer runs
        extra['sequential_row_sparse'] = sequential_row_sparse(wl, min(K, 20), B)
        # SURVEY 8 D2 variants: tables drawn with std 0.1 (losses away from ln 2) and Zipf(1.05) item popularity (hot rows)
        del wl.gu, wl.gi
        variants = {}
        for name, kw in (('std0.1', dict(std=0.1)), ('zipf1.05_std0.1', dict(std=0.1, zipf=1.05)),
                         ('zipf1.05_std0.1_hot_rows', dict(std=0.1, zipf=1.05, hot_rows=True))):
            v = RecWorkload(scale, B, K, W, Re, dev, grad_mode=args.grad_mode, **kw)
            ms_v, lm = v.timed(timer)
            variants[name] = dict(rate_fields(ms_v, K, B, 1, peak), loss_mean=lm)
            if v.hot is not None:
                variants[name]['hot_rows'] = {'users': int(v.hot[0].numel()), 'items': int(v.hot[1].numel()),
                                              'note': 'rows named by >= 0.2 % of the ids: gradients pre-aggregated per CTA in shared memory'}
            del v
            torch.cuda.empty_cache()
        extra['variants'] = variants
        wl.gu, wl.gi = torch.zeros_like(wl.ut), torch.zeros_like(wl.it)

    # ---- end-to-end through the public trainer API, host batches ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        from recbole_cdr_b200.trainer import FusedStepRunner
        chunk = args.chunk if args.chunk > 0 else (max(1, K // 2) if K < 100 else 50)
        n_chunks = max(1, K // chunk)
        Ke = chunk * n_chunks
        if sharded:
            def launch(idb, _label, o8):
                sh_runner.run(idb, out8=o8)
            runner = FusedStepRunner({'pairwise': True}, launch=launch, device=dev, n_buffers=3)
            src_ids = ids
        else:
            spec = dict(user_tab=wl.ut, item_tab=wl.it, pairwise=True, reg_weight=0.01, gamma=1e-10)
            runner = FusedStepRunner(spec, lr=None, grad_tables=(wl.gu, wl.gi), n_buffers=3)
            if args.grad_mode == 'fresh':   # the K-step block's gradient: the map is cleared per block, chunks accumulate
                runner.steps_kw = dict(touch=wl.touch_map(), fresh=False)
            src_ids = wl.ids
        Re2 = max(3, R)   # (as many samples as the device-timed value: the median must survive host hiccups -- the clock sampler's
                          # nvidia-smi fork, a first cudaHostAlloc -- that land inside a 0.2 ms pass)
        host = src_ids[W: W + min(R, Re2) * K].cpu().pin_memory()   # pinned [*, 3, B] int64 id blocks on the host
        # untimed passes of exactly the timed code path: buffers of every shape created, and the pinned-memory pool filled with
        # the loss tensors of three passes (a first-time cudaHostAlloc inside the timed region cost 0.4 ms of a 0.2 ms pass)
        warm = []
        for _ in range(3):
            warm.append([runner.run(host[c * chunk:(c + 1) * chunk]) for c in range(n_chunks)])
            timer.barrier()
        del warm
        api_ms, api_loss = [], []
        import gc
        gc.collect()
        gc.disable()   # (no cyclic-GC pause inside a 0.2 ms host-driven pass; re-enabled below)
        for r in range(min(R, Re2)):
            blocks = [host[r * K + c * chunk: r * K + (c + 1) * chunk] for c in range(n_chunks)]
            timer.barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            if not sharded and args.grad_mode == 'fresh':
                wl.touch_map().clear()
            losses = [runner.run(b) for b in blocks]   # per chunk: H2D ids -> one persistent launch -> D2H losses
            a1.record()
            timer.barrier()
            api_ms.append(a0.elapsed_time(a1))
            api_loss.append(float(torch.stack([l.mean() for l in losses]).mean()))
        gc.enable()
        t = torch.tensor(api_ms, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        api_ms = t.tolist()
        med = statistics.median(api_ms)
        e2e = {'value': world * B * Ke / (med * 1e-3), 'unit': 'interactions/s',
               'h2d_bytes_per_step': 3 * 8 * B, 'd2h_bytes_per_step': 32, 'steps': Ke, 'steps_per_launch': chunk,
               'repeats': len(api_ms), 'min_ms': min(api_ms), 'max_ms': max(api_ms), 'samples_ms': [round(x, 4) for x in api_ms],
               'loss_mean': statistics.mean(api_loss),
               'api': 'trainer.FusedStepRunner.run(pinned [chunk,3,B] int64 ids): H2D copy -> xdr_train_steps -> D2H losses, '
                      'chunks triple-buffered so copies overlap launches; host enqueue time included'}
    clk = clocks.stop() if rank == 0 else None

    if rank == 0:
        med = statistics.median(ms_list)
        value = world * B * K / (med * 1e-3)
        achieved = BYTES_PER_INTERACTION_BPR_D64 * B * K / (med * 1e-3) / 1e9   # per GPU
        traffic, traffic_note = ncu_traffic(B, K, sharded, args.grad_mode)
        line = {
            'metric': 'interactions/sec (gather+map+score+scatter)', 'value': value, 'unit': 'interactions/s',
            'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': med / K, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(workload, B, world, ds, {'grad_mode': args.grad_mode}),
            'timing': Timer.summary(ms_list),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'traffic_note': traffic_note, 'peak_source': peak_src, 'kernel': kernel_name,
                         'units_per_launch': B * K, 'bytes_per_interaction': BYTES_PER_INTERACTION_BPR_D64},
            'e2e': e2e, 'gpu_launches': launches, 'clocks': clk, 'loss_mean': loss_mean,
        }
        line.update(extra)
    # ---- N = 1: the 10M x 10M tables on one GPU (denominator of the 8-GPU claim of BASELINE.json's north_star) -------
    if not sharded and not args.no_extras and workload == 'emcdr_1m':
        try:
            del wl
            torch.cuda.empty_cache()
            big = RecWorkload(SCALES['emcdr_10m'], B, K, W, max(3, R // 2), dev, grad_mode=args.grad_mode)
            ms_b, lm = big.timed(timer)
            line['emcdr_10m'] = dict(rate_fields(ms_b, K, B, 1, peak), loss_mean=lm, users_total=big.ds.num_total_user,
                                     items_total=big.ds.num_total_item,
                                     note='BASELINE configs[4] tables (17.9 GB incl. gradients per domain pair) on ONE GPU, '
                                          'B = 8192: the 1-GPU denominator of the row-sharded N > 1 runs')
            del big
            torch.cuda.empty_cache()
        except Exception as e:  # never lose the main line to the sub-run
            line['emcdr_10m'] = {'error': repr(e)[:200]}
        # ---- the dense rows of the path, so that the default line carries them too: the OVERLAP-phase map step ("map" of the
        # metric's name) and CoNet's BOTH step at BASELINE configs[2], each measured by THIS script in a child process (its own
        # CUDA context: nothing of it can disturb the numbers above, which are final by now)
        try:
            line['model_steps'] = model_step_sublines()
        except Exception as e:
            line['model_steps'] = {'error': repr(e)[:200]}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            r = cpu_reference_rates(scale, B, 1, args.cpu_steps)
            line['cpu_baseline'] = cpu_baseline_object(r, B, scale, args.cpu_steps)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def model_step_sublines():
    """`bench.py --workload emcdr_map` and `--workload conet_5m` as child processes; a summary of each line (or its error)."""
    import subprocess
    out = {}
    for name, flags in (('emcdr_map', ['--steps', '20', '--warmup', '5']),
                        ('conet_5m', ['--steps', '10', '--warmup', '3', '--repeats', '5'])):
        cmd = [sys.executable, os.path.abspath(__file__), '--workload', name, '--no-cpu-baseline'] + flags
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=150)
            rows = [l for l in p.stdout.splitlines() if l.startswith('{')]
            if p.returncode != 0 or not rows:
                out[name] = {'error': ('rc=%d ' % p.returncode) + p.stderr[-300:]}
                continue
            d = json.loads(rows[-1])
            out[name] = {'workload': d['config']['workload'], 'engine': d['config'].get('engine'), 'value': d['value'],
                         'unit': d['unit'], 'ms_per_step': d['ms_per_step'], 'steps': d['steps'], 'timing': d.get('timing'),
                         'roofline_frac': d['roofline']['frac'], 'bytes_per_interaction': d['roofline']['bytes_per_interaction'],
                         'e2e': d.get('e2e'), 'gpu_launches': d.get('gpu_launches'), 'loss_mean': d.get('loss_mean'),
                         'dtype': d.get('dtype'), 'command': 'python bench.py ' + ' '.join(cmd[2:])}
        except Exception as e:   # a time-out or an unreadable line: the main line goes out without this row
            out[name] = {'error': repr(e)[:300]}
    return out


def ncu_traffic(B, K, sharded, grad_mode):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this kernel
    (profiles/), scaled from the captured launch's step count to K; None when no capture matches the configuration."""
    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if sharded or B != 8192 or not os.path.exists(path):
        return None, 'no ncu capture for this configuration'
    try:
        rec = json.load(open(path)).get(grad_mode)
        if not rec:
            return None, 'no ncu capture for this gradient mode'
        return rec['dram_bytes_per_step'] * K, rec['note']
    except Exception:
        return None, 'profiles/ncu_traffic.json unreadable'


def per_step_comparison(wl, K, W, B, peak):
    """The per-step fwd/bwd kernel pair (strictly sequential reference semantics) replayed from a CUDA graph, and eagerly."""
    from recbole_cdr_b200 import _lib
    ut, it, gu, gi, ids = wl.ut, wl.it, wl.gu, wl.gi, wl.ids
    scores = torch.empty(2, B, device=wl.dev)
    out8 = torch.empty(K + W, 8, device=wl.dev)
    ws = _lib.workspace(wl.dev)

    def step(s):
        u, ip, ineg = ids[s, 0], ids[s, 1], ids[s, 2]
        _lib.call('xdr_bpr_fwd', ut.data_ptr(), it.data_ptr(), ut.shape[0], it.shape[0], D, u.data_ptr(), ip.data_ptr(),
                  ineg.data_ptr(), B, 1e-10, 0.01, scores[0].data_ptr(), scores[1].data_ptr(), out8[s].data_ptr(),
                  ws.data_ptr(), None, _lib.cur_stream())
        _lib.call('xdr_bpr_bwd', ut.data_ptr(), it.data_ptr(), ut.shape[0], it.shape[0], D, u.data_ptr(), ip.data_ptr(),
                  ineg.data_ptr(), B, 1e-10, 0.01, scores[0].data_ptr(), scores[1].data_ptr(), out8[s].data_ptr(), None,
                  1.0, gu.data_ptr(), gi.data_ptr(), _lib.cur_stream())

    out = {}
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
    out['per_step_kernels_cuda_graph'] = {'ms_per_step': gms / K, 'value': B * K / (gms * 1e-3), 'launches': 2 * K,
                                          'roofline_frac': BYTES_PER_INTERACTION_BPR_D64 * B * K / (gms * 1e-3) / 1e9 / peak}
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for s in range(W, W + K):
        step(s)
    p1.record()
    torch.cuda.synchronize()
    pms = p0.elapsed_time(p1)
    out['per_step_kernels_eager'] = {'ms_per_step': pms / K, 'value': B * K / (pms * 1e-3), 'launches': 2 * K}
    return out


def _import_reference_model(name):
    """An unmodified reference model class from oracle/_ref (staged by build()) over the recbole stub; None if not staged."""
    if _import_reference_emcdr() is None:
        return None
    import importlib
    mod = importlib.import_module('recbole_cdr.model.cross_domain_recommender.' + name.lower())
    return getattr(mod, name)


MODEL_WORKLOADS = {
    # name: (model class, config overrides, bytes per interaction, what a step is)
    'emcdr_map': ('EMCDR', 1032, 'EMCDR OVERLAP-phase map step (gather -> MLP 64-128-64 -> MSE -> backward -> scatter), synthetic '
                                 'emcdr_1m tables, b = %d overlapped users per step'),
    'conet_5m': ('CoNet', 4116, 'CoNet BOTH-phase step (BASELINE configs[2]: 5M users / 2M items, dim 128, cross-stitch MLP '
                                '[256, 64, 32, 16, 8]): source + target tower passes on %d rows per domain, fwd + bwd + scatter'),
}


def _model_workload(name, batch, device):
    """(dataset, config, batch factory, interactions per step) of a model-step workload, on `device`."""
    synthetic = load_synthetic() if str(device) == 'cpu' else None
    if synthetic is None:
        from recbole_cdr_b200.data import synthetic
    base = {'source_domain': {'NEG_PREFIX': 'neg_'}, 'target_domain': {'NEG_PREFIX': 'neg_'}, 'device': device}
    if name == 'emcdr_map':
        ds = synthetic.emcdr_scale(SCALES['emcdr_1m'])
        cfg = dict(base, latent_factor_model='BPR', source_embedding_size=D, target_embedding_size=D, reg_weight=0.01,
                   mapping_function='non_linear', mlp_hidden_size=[128], overlap_batch_size=batch)

        def make(seed):
            g = torch.Generator(device=device).manual_seed(seed)
            return {'overlap': torch.randint(0, ds.num_overlap_user, (batch, 1), device=device, generator=g)}
        return ds, cfg, make, batch, 'OVERLAP'
    ds = synthetic.SyntheticCrossDomainDataset(synthetic.IdSpace(2_500_001, 2_500_000, 2_500_000),
                                               synthetic.IdSpace(1, 2_000_000, 2_000_000))
    cfg = dict(base, embedding_size=128, reg_weight=0.01, mlp_hidden_size=[64, 32, 16, 8])

    def make(seed):
        b = synthetic.make_batch(ds, 'source', batch, 2 * seed, device, pairwise=False)
        b.update(synthetic.make_batch(ds, 'target', batch, 2 * seed + 1, device, pairwise=False))
        return b
    return ds, cfg, make, 2 * batch, None


def cpu_reference_model_rate(name, batch, n_timed):
    """The reference model class's step (calculate_loss + backward) on the host cores."""
    cls = _import_reference_model(MODEL_WORKLOADS[name][0])
    if cls is None:
        return None
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ds, cfg, make, units, phase = _model_workload(name, batch, 'cpu')
    torch.manual_seed(2022)
    model = cls(cfg, ds)
    if phase:
        model.set_phase(phase)
    ts = []
    for s_ in range(1 + n_timed):
        inter = make(100 + s_)
        model.zero_grad(set_to_none=True)
        t0 = time.perf_counter()
        loss = model.calculate_loss(inter)
        (sum(loss) if isinstance(loss, tuple) else loss).sum().backward()
        dt = time.perf_counter() - t0
        if s_ >= 1:
            ts.append(dt)
    sec = sum(ts) / len(ts)
    return {'value': units / sec, 'unit': 'interactions/s', 'cores': cores, 'kind': 'reference',
            'sample': f'{n_timed} timed + 1 warm-up steps of the same workload on the unmodified reference {MODEL_WORKLOADS[name][0]} '
                      f'class (oracle/_ref), forward + backward to dense grads, torch CPU {cores} threads, {sec * 1e3:.0f} ms/step'}


def run_model_step(args, dev, name):
    """--workload emcdr_map | conet_5m: one training step of a drop-in model class (calculate_loss + backward: gather -> dense
    map / cross-stitch layers on tcgen05 -> loss -> backward -> scatter-add into the tables' gradient rows), captured as a CUDA
    graph (trainer.GraphedTrainStep).  A step = copy of the step's ids into the graph's input buffers + one replay."""
    add_paths()
    from recbole_cdr_b200 import _lib
    from recbole_cdr_b200.data import Interaction
    from recbole_cdr_b200.trainer import GraphedTrainStep
    import importlib
    cls_name, bytes_per, what = MODEL_WORKLOADS[name]
    cls = getattr(importlib.import_module('recbole_cdr_b200.model.cross_domain_recommender.' + cls_name.lower()), cls_name)
    b = args.batch if name == 'emcdr_map' else (16384 if args.batch == 8192 else args.batch)
    K, W, R = args.steps, max(3, args.warmup), max(1, args.repeats)
    peak, peak_src = measured_peaks()
    if args.dense_engine >= 0:
        _lib._lib.xdr_set_dense_engine(int(args.dense_engine))
    from recbole_cdr_b200 import ops as _ops
    _ops.set_table_grad_mode('inplace')   # table gradients are scatter-added into persistent .grad buffers (no dense [N, D] temporaries)
    ds, cfg, make, units, phase = _model_workload(name, b, dev)
    if name == 'emcdr_map':
        if args.map_engine:   # default: the model's own choice ('auto': the tcgen05 map-step kernel where it applies)
            cfg['xdr_fused_mlp'] = False if args.map_engine == 'composed' else args.map_engine
    if args.dense_engine >= 0:
        cfg['xdr_dense_engine'] = int(args.dense_engine)   # (models that choose their layers' engine themselves: CoNet)
    torch.manual_seed(2022)
    with torch.device(dev):
        model = cls(cfg, ds)
    if phase:
        model.set_phase(phase)
    dense_engine = getattr(model, 'dense_engine', max(0, args.dense_engine))
    n = W + R * K
    batches = [Interaction(make(s_)) for s_ in range(n)]
    # how many libxdr entry points one step goes through (each launches at least one kernel of this repository)
    counted, real_call = [0], _ops.call

    def counting_call(nm, *a_, **k_):
        counted[0] += 1
        return real_call(nm, *a_, **k_)
    _ops.call = counting_call
    try:   # (counted over the 3 warm-up steps + the captured one inside GraphedTrainStep)
        step = GraphedTrainStep(model, batches[0])
    finally:
        _ops.call = real_call
    calls_per_step = counted[0] // 4
    timer = Timer(dev, 1)
    clocks = ClockSampler(dev.index or 0)
    clocks.start()

    def run_steps(lo, hi):
        for k in range(lo, hi):
            step(batches[k])

    run_steps(0, W)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_steps(0, K)
    host_s = time.perf_counter() - t0          # host enqueue time of K steps: the gate must outlast it
    torch.cuda.synchronize()
    global GATE_CYCLES
    GATE_CYCLES = int(max(GATE_CYCLES, 2.5 * host_s * 1.965e9))
    ms_list = timer.samples(lambda r: run_steps(W + r * K, W + (r + 1) * K), R)
    loss = float(step.loss)
    med = statistics.median(ms_list)
    # end to end: ids (and labels) from pinned host memory, loss back to the host, every step
    host_b = [Interaction({k: batches[W + i][k].cpu().pin_memory() for k in batches[W + i].columns}) for i in range(K)]
    h2d = sum(int(host_b[0][k].numel() * host_b[0][k].element_size()) for k in host_b[0].columns)
    losses = torch.empty(K, dtype=torch.float32, pin_memory=True)
    api = []
    for r in range(max(3, R // 2)):
        timer.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for k in range(K):
            l = step(host_b[k])   # H2D copy of the step's fields into the graph's input buffers + replay
            losses[k].copy_(l, non_blocking=True)
        a1.record()
        timer.barrier()
        api.append(a0.elapsed_time(a1))
    clk = clocks.stop()
    amed = statistics.median(api)
    achieved = bytes_per * units * K / (med * 1e-3) / 1e9
    engine = {0: 'fp32 FMA dense kernels', 1: 'tcgen05 dense engine (tc5_dense.cu)',
              2: 'tcgen05 (tc5_dense.cu) for forward / input gradient of the wide layers, fp32 FMA for the rest'}[dense_engine]
    if name == 'emcdr_map':
        engine = {'tc5': 'tc5_mlp_kernel (one fused tcgen05 kernel per pass)', 'tc': 'tc_mlp_kernel (mma.sync)',
                  'fma': 'fused_mlp_kernel (fp32)', 'composed': 'composed fp32 FMA kernels'}.get(args.map_engine or 'tc5', args.map_engine)
    line = {
        'metric': 'interactions/sec (gather+map+score+scatter)', 'value': units * K / (med * 1e-3), 'unit': 'interactions/s',
        'n_gpus': 1, 'steps': K, 'warmup': W, 'ms_per_step': med / K, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None,
        'dtype': 'f32 (dense products: bf16x3 on tcgen05, fp32 accumulate)' if (dense_engine or (name == 'emcdr_map' and args.map_engine in ('', 'tc5'))) else 'f32',
        'data': 'synthetic',
        'config': {'workload': what % b, 'engine': engine, 'users_total': ds.num_total_user, 'items_total': ds.num_total_item,
                   'batch_per_gpu': b, 'l2': 'inputs larger than L2 (tables of 0.8 .. 7 GB, uniform random rows)',
                   'parallelism': 'single GPU',
                   'graph_lanes': {-1: 2, 0: 1, 1: 1}.get(_ops.CROSS_STREAMS, _ops.CROSS_STREAMS) if name == 'conet_5m' else 1},
        'timing': Timer.summary(ms_list),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                     'peak_source': peak_src, 'kernel': engine, 'units_per_launch': units, 'bytes_per_interaction': bytes_per,
                     'note': 'a step is a CUDA graph of %d library calls; at these batch sizes it is bound by launch gaps and '
                             'dependent-kernel latency, not by HBM (the algorithmic bytes of a step are %.1f MB)' %
                             (calls_per_step, bytes_per * units / 1e6)},
        'e2e': {'value': units * K / (amed * 1e-3), 'unit': 'interactions/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                'steps': K, 'repeats': len(api), 'min_ms': min(api), 'max_ms': max(api),
                'api': f'{cls_name}.calculate_loss + backward replayed by trainer.GraphedTrainStep: per step H2D fields -> graph -> D2H loss'},
        'gpu_launches': calls_per_step * K, 'clocks': clk, 'loss_mean': loss,
    }
    if not args.no_cpu_baseline:
        cb = cpu_reference_model_rate(name, b, args.cpu_steps)
        if cb:
            line['cpu_baseline'] = cb
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_xdr(args)


if __name__ == '__main__':
    main()
