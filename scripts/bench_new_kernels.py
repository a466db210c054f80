"""Micro-benchmarks of the kernels written while no GPU was reachable (tc_mlp.cu, tc_conet.cu, sparse_optim.cu,
topk_score.cu), each next to the path it is meant to replace, at the BASELINE config shapes (one GPU).

Every section first checks the new kernel against the composed / library path on the same inputs (loss within 1e-4) and
only then times it; a failing section prints {"kernel": ..., "error": ...} and the script moves on.  CUDA events, best of 5
after warm-up; algorithmic bytes / FLOPs are SURVEY.md section 8 D3's.  Output: one JSON line per row (also appended to
gpurun_out/new_kernels.jsonl)."""
import json, os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'recbole-cdr_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from fake_data import base_config
from recbole_cdr_b200 import _lib, ops
from recbole_cdr_b200.data import Interaction, synthetic
from recbole_cdr_b200.data.idspace import IdSpace
from recbole_cdr_b200.data.synthetic import SyntheticCrossDomainDataset
from recbole_cdr_b200.trainer import GraphedTrainStep

dev = torch.device('cuda', 0)
SMALL = os.environ.get('XDR_SMALL') == '1'
SKIP_CHECK = os.environ.get('XDR_SKIP_CHECK') == '1'     # diagnostic engines (one TF32 pass) are not parity-grade
PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
OUT = open(os.path.join(ROOT, 'gpurun_out', 'new_kernels.jsonl'), 'a')


def timeit(fn, reps=5, inner=1):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best * 1e-3


def report(name, sec, bytes_=None, flops=None, units=None, unit_name='inter', **extra):
    line = {'kernel': name, 'us': round(sec * 1e6, 2), 'lib': os.path.basename(_lib.LIB_PATH)}
    if bytes_:
        line.update(GBps=round(bytes_ / sec / 1e9, 1), hbm_frac=round(bytes_ / sec / 1e9 / PEAK, 4))
    if flops:
        line['TFLOPs'] = round(flops / sec / 1e12, 2)
    if units:
        line[f'M{unit_name}_per_s'] = round(units / sec / 1e6, 1)
    line.update(extra)
    print(json.dumps(line), flush=True)
    OUT.write(json.dumps(line) + '\n'); OUT.flush()


def section(fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001 -- keep going: every section is independent
        line = {'kernel': fn.__name__, 'error': f'{type(e).__name__}: {e}', 'trace': traceback.format_exc()[-600:]}
        print(json.dumps(line), flush=True)
        OUT.write(json.dumps(line) + '\n'); OUT.flush()
    torch.cuda.synchronize()


def both_batch(dsx, Bx, seed):
    b = synthetic.make_batch(dsx, 'source', Bx, seed, dev, pairwise=False)
    b.update(synthetic.make_batch(dsx, 'target', Bx, seed + 1, dev, pairwise=False))
    return Interaction(b)


ops.set_table_grad_mode('inplace')
g = torch.Generator(device=dev).manual_seed(0)


def emcdr_map_step():
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    ds = synthetic.emcdr_scale(1_000_000 if not SMALL else 10_000)
    b = 8192
    ov = torch.randint(0, ds.num_overlap_user, (b, 1), device=dev, generator=g)
    losses = {}
    # the tcgen05 engine only on request (XDR_BENCH_TC5=1, in its own process): a wrong descriptor reading would trap the context
    engines = (False, 'fma', 'tc') + (('tc5',) if os.environ.get('XDR_BENCH_TC5', '0') == '1' else ())
    for engine in engines:
        cfg = base_config(device=dev, latent_factor_model='BPR', source_embedding_size=64, target_embedding_size=64,
                          reg_weight=0.01, mapping_function='non_linear', mlp_hidden_size=[128], xdr_fused_mlp=engine)
        torch.manual_seed(1)
        with torch.device(dev):
            m = EMCDR(cfg, ds)
        m.set_phase('OVERLAP')
        inter = Interaction({'overlap': ov})
        losses[engine] = float(m.calculate_loss(inter))
        def step():
            m.calculate_loss(inter).backward()
        name = {False: 'composed kernels', 'fma': 'fp32 row-tile kernel', 'tc': 'tensor-core row-tile kernel (NEW)',
                'tc5': 'tcgen05 kernel, TMEM-resident weight gradients (NEW)'}[engine]
        report(f'A4 EMCDR map step fwd+bwd, {name}, eager, b = 8192', timeit(step), bytes_=b * 1032, flops=b * 98304, units=b)
        gs = GraphedTrainStep(m, inter)
        report(f'A4 EMCDR map step fwd+bwd, {name}, CUDA-graph replay, b = 8192', timeit(lambda: gs(inter), inner=10),
               bytes_=b * 1032, flops=b * 98304, units=b)
        del m, gs
    for engine in engines[2:]:
        assert SKIP_CHECK or abs(losses[engine] - losses[False]) <= 1e-4 * abs(losses[False]), losses


def dtcdr_both_step():
    from recbole_cdr_b200.model.cross_domain_recommender.dtcdr import DTCDR
    s = 1 if not SMALL else 100
    dsb = SyntheticCrossDomainDataset(IdSpace(500_001 // s, 500_000 // s, 500_000 // s), IdSpace(200_001 // s, 400_000 // s, 400_000 // s))
    ib = both_batch(dsb, 8192, 5)
    losses = {}
    for engine in (False, 'fma', 'tc'):
        torch.manual_seed(1)
        with torch.device(dev):
            m = DTCDR(base_config(device=dev, embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF',
                                  alpha=0.5, xdr_fused_mlp=engine), dsb)
        losses[engine] = float(m.calculate_loss(ib))
        def step():
            m.calculate_loss(ib).backward()
        name = {False: 'composed kernels', 'fma': 'fp32 row-tile kernel', 'tc': 'tensor-core row-tile kernel (NEW)'}[engine]
        report(f'A14-15 DTCDR BOTH step fwd+bwd, {name}, eager, 2 x B=8192', timeit(step), bytes_=2 * 8192 * 2068,
               flops=2 * 8192 * 27700, units=2 * 8192)
        gs = GraphedTrainStep(m, ib)
        report(f'A14-15 DTCDR BOTH step fwd+bwd, {name}, CUDA-graph replay', timeit(lambda: gs(ib), inner=10),
               bytes_=2 * 8192 * 2068, flops=2 * 8192 * 27700, units=2 * 8192)
        del m, gs
    assert abs(losses['tc'] - losses[False]) <= 1e-4 * abs(losses[False]), losses


def conet_both_step():
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    s = 1 if not SMALL else 100
    dsc = SyntheticCrossDomainDataset(IdSpace(2_500_001 // s, 2_500_000 // s, 2_500_000 // s), IdSpace(1, 2_000_000 // s, 2_000_000 // s))
    ic = both_batch(dsc, 16384, 9)
    losses = {}
    for fused in (False, True):
        torch.manual_seed(1)
        with torch.device(dev):
            m = CoNet(base_config(device=dev, embedding_size=128, reg_weight=0.01, mlp_hidden_size=[64, 32, 16, 8],
                                  xdr_fused_conet=fused), dsc)
        losses[fused] = float(m.calculate_loss(ic))
        def step():
            m.calculate_loss(ic).backward()
        name = 'ONE tensor-core kernel per tower pass (NEW)' if fused else 'composed kernels'
        report(f'A7-8 CoNet BOTH step fwd+bwd, {name}, eager, config #3 shape, 2 x B=16384', timeit(step, reps=3),
               bytes_=2 * 16384 * 4116, flops=2 * 16384 * 458000, units=2 * 16384)
        gs = GraphedTrainStep(m, ic)
        report(f'A7-8 CoNet BOTH step fwd+bwd, {name}, CUDA-graph replay', timeit(lambda: gs(ic), reps=3, inner=5),
               bytes_=2 * 16384 * 4116, flops=2 * 16384 * 458000, units=2 * 16384)
        del m, gs
    assert SKIP_CHECK or abs(losses[True] - losses[False]) <= 1e-4 * abs(losses[False]), losses


def sparse_optimizers():
    N, D, n = (2_000_001, 64, 3 * 8192) if not SMALL else (20_001, 64, 3 * 8192)
    w = torch.randn(N, D, device=dev) * 0.01
    grad = torch.zeros_like(w)
    ids = torch.randint(0, N, (n,), device=dev, generator=g)
    rows = torch.randn(n, D, device=dev) * 0.01
    stamp = torch.zeros(N, dtype=torch.int32, device=dev)
    s1, s2 = torch.zeros_like(w), torch.zeros_like(w)
    step = [0]
    for kind, nm, rw in ((_lib.OPT_SGD, 'SGD', 4), (_lib.OPT_ADAGRAD, 'Adagrad', 6), (_lib.OPT_LAZY_ADAM, 'lazy Adam', 8)):
        def fn():
            step[0] += 1
            ops.scatter_add_rows_raw(grad, ids, rows)
            ops.sparse_optim_rows(kind, w, grad, ids, stamp, step[0], 1e-3, state1=s1 if kind else None,
                                  state2=s2 if kind == _lib.OPT_LAZY_ADAM else None, adam_t=step[0])
        t_both = timeit(fn)
        t_sc = timeit(lambda: ops.scatter_add_rows_raw(grad, ids, rows))
        grad.zero_()
        report(f'F1 row-sparse {nm} over {n} ids (scatter-add excluded), 2M x 64 tables', max(t_both - t_sc, 1e-9),
               bytes_=n * (12 + rw * 256), units=n, unit_name='rows')
    # what it replaces: dense torch.optim.Adam over the same table
    p = torch.nn.Parameter(w.clone())
    opt = torch.optim.Adam([p], lr=1e-3)
    p.grad = torch.zeros_like(p)
    report('F1 reference: dense torch.optim.Adam step over the 2M x 64 table (+ dense zero_grad)',
           timeit(lambda: (opt.step(), p.grad.zero_())), bytes_=N * D * 4 * 8, units=n, unit_name='rows')


def full_sort_topk():
    for B, n_items in ((4096, 1_000_001), (256, 1_000_001)) if not SMALL else ((256, 20_001),):
        D, k = 64, 20
        U = torch.randn(B, D, device=dev) * 0.1
        I = torch.randn(n_items, D, device=dev) * 0.1
        lens = torch.randint(0, 30, (B,), device=dev, generator=g)
        hp = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(lens, 0)])
        hi = torch.sort(torch.randint(1, n_items, (int(hp[-1]),), device=dev, generator=g).view(-1))[0]
        # per-user ascending order: sort inside each segment
        seg = torch.repeat_interleave(torch.arange(B, device=dev), lens)
        order = torch.argsort(seg * n_items + hi)
        hi = hi[order]
        sc, pos = ops.full_sort_topk(U, I, k, hist_ptr=hp, hist_ids=hi)
        def ref():
            full = U @ I.T
            full[:, 0] = -float('inf')
            full[seg, hi] = -float('inf')
            return torch.topk(full, k, dim=1)
        rs, ri = ref()
        if not SKIP_CHECK:
            torch.testing.assert_close(sc, rs, rtol=2e-5, atol=1e-6)
        fl = 2.0 * B * n_items * D
        report(f'F2 fused full-sort top-{k} (NEW), {B} users x {n_items} items, D = 64', timeit(lambda: ops.full_sort_topk(U, I, k, hist_ptr=hp, hist_ids=hi), reps=3),
               bytes_=n_items * D * 4 * ((B + 63) // 64) + B * D * 4, flops=fl, units=B, unit_name='users')
        try:
            sc5, _ = ops.full_sort_topk(U, I, k, hist_ptr=hp, hist_ids=hi, engine='tc5')
            if not SKIP_CHECK:
                torch.testing.assert_close(sc5, rs, rtol=2e-5, atol=1e-6)
            report(f'F2 fused full-sort top-{k}, tcgen05 engine (NEW), {B} users x {n_items} items, D = 64',
                   timeit(lambda: ops.full_sort_topk(U, I, k, hist_ptr=hp, hist_ids=hi, engine='tc5'), reps=3),
                   bytes_=n_items * D * 4 * ((B + 127) // 128) + B * D * 4, flops=fl, units=B, unit_name='users')
        except Exception as e:  # noqa: BLE001 -- the descriptor reading may need the correction ubench_tcgen05 reports
            print(json.dumps({'kernel': 'F2 tcgen05 engine', 'error': f'{type(e).__name__}: {str(e)[:300]}'}), flush=True)
        report(f'F2 reference: torch.matmul + mask + torch.topk, {B} users x {n_items} items', timeit(ref, reps=3),
               bytes_=B * n_items * 4 * 3, flops=fl, units=B, unit_name='users')


def early_scatter_steps():
    """reg_weight == 0 launches of the persistent kernel with and without the early-scatter form (steps_persistent.cu EARLY):
    (a) the headline shape (1M x 1M rows, dim 64, B = 8192 BPR triples, K = 200 steps), (b) CMF at BASELINE configs[0]'s sizes
    (6982 users x 3351 items, B = 2048 pointwise rows, lambda = gamma = 0 as in CMF.yaml)."""
    cases = [('EMCDR BPR reg_weight 0, 1M x 1M, B = 8192', 1_000_001 if not SMALL else 10_001, 1_000_001 if not SMALL else 10_001, 8192, True, 1560),
             ('CMF term lambda 0, 6982 x 3351 (bundled sizes), B = 2048', 6982, 3351, 2048, False, 1044)]
    K = 200
    for name, nu, ni, B, pairwise, bytes_per in cases:
        ut, it = torch.randn(nu, 64, device=dev) * 0.1, torch.randn(ni, 64, device=dev) * 0.1
        gu, gi = torch.zeros_like(ut), torch.zeros_like(it)
        u = torch.randint(1, nu, (K, B), device=dev, generator=g)
        ia, ib = torch.randint(1, ni, (K, B), device=dev, generator=g), torch.randint(1, ni, (K, B), device=dev, generator=g)
        y = (torch.rand(K, B, device=dev, generator=g) < 0.5).float()
        losses = {}
        for early in (False, True):
            ops.set_steps_early_scatter(early)
            try:
                def step():
                    if pairwise:
                        return ops.train_steps(ut, it, u, ia, ib, reg_weight=0.0, user_dst=gu, item_dst=gi)[0]
                    return ops.train_steps(ut, it, u, ia, None, y, loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=0.0, user_dst=gu,
                                           item_dst=gi)[0]
                losses[early] = step()[:, 0].clone()
                t = timeit(step)
            finally:
                ops.set_steps_early_scatter(False)
            report(f'A3/A16 {name}, K = {K} steps per launch, {"early scatter (NEW)" if early else "default kernel"} (per step)',
                   t / K, bytes_=B * bytes_per, units=B)
        assert SKIP_CHECK or torch.equal(losses[False], losses[True]), 'early scatter changed the per-step losses'


ALL = (emcdr_map_step, dtcdr_both_step, conet_both_step, sparse_optimizers, full_sort_topk, early_scatter_steps)
WANT = os.environ.get('XDR_SECTIONS')                      # comma-separated subset of the section names
for sec in ALL:
    if WANT is None or sec.__name__ in WANT.split(','):
        section(sec)
