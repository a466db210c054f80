"""Per-row micro-benchmarks of the SURVEY section 8 hot-path kernels at the BASELINE config shapes (one GPU).

Prints one line per kernel group: time per call (CUDA events, best of 5 after warm-up), algorithmic bytes / FLOPs
(SURVEY 8 D3 formulas), achieved GB/s or TFLOP/s and the fraction of the measured peaks.  Not the driver's bench line
(that is bench.py); this is the evidence for the rows other than the headline."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'recbole-cdr_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from recbole_cdr_b200 import _lib, ops
from recbole_cdr_b200.data import Interaction, synthetic
from recbole_cdr_b200.data.idspace import IdSpace

dev = torch.device('cuda', 0)
ONLY = os.environ.get('XDR_ROWS')  # e.g. 'models' to run only the composed-model steps
PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0
results = []


def timeit(fn, reps=5, inner=1):
    for _ in range(2):
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


def report(name, sec, bytes_=None, flops=None, units=None, unit_name='rows'):
    line = {'kernel': name, 'us': sec * 1e6}
    if bytes_:
        line.update(GBps=bytes_ / sec / 1e9, hbm_frac=bytes_ / sec / 1e9 / PEAK)
    if flops:
        line['TFLOPs'] = flops / sec / 1e12
    if units:
        line[f'M{unit_name}_per_s'] = units / sec / 1e6
    results.append(line)
    print(json.dumps(line))


g = torch.Generator(device=dev).manual_seed(0)

# ---- A1: gather / scatter-add, 2M x 64 table, 1M random rows ------------------------------------------------------
N, D, n = (2_000_001, 64, 1 << 20) if not ONLY else (1001, 64, 1024)
tab = torch.randn(N, D, device=dev) * 0.01
idx = torch.randint(0, N, (n,), device=dev, generator=g)
out = torch.empty(n, D, device=dev)
report('A1 gather_rows 1M rows x 256 B', timeit(lambda: ops.gather_rows_raw(tab, idx, out)), bytes_=n * (8 + 2 * 256), units=n)
grad = torch.zeros_like(tab)
report('A1 scatter_add_rows 1M rows x 256 B', timeit(lambda: ops.scatter_add_rows_raw(grad, idx, out)), bytes_=n * (8 + 2 * 256), units=n)

# ---- A3/A16: pointwise persistent steps (EMCDR-MF / CMF domain term), B = 8192 (config #2 shape), K = 64 -----------
ds = synthetic.emcdr_scale(1_000_000)
ut = torch.randn(ds.num_total_user, D, device=dev) * 0.01
it = torch.randn(ds.num_total_item, D, device=dev) * 0.01
gu, gi = torch.zeros_like(ut), torch.zeros_like(it)
K, B = (64, 8192) if not ONLY else (2, 8192)
def batches(pairwise):
    bs = [synthetic.make_batch(ds, 'source', B, 1 + s, 'cpu', pairwise=pairwise) for s in range(K)]
    return {k: torch.stack([b[k] for b in bs]).to(dev) for k in bs[0]}
bp = batches(False)
lab = bp['source_label'].contiguous()
for kind, nm in ((_lib.LOSS_MSE, 'EMCDR-MF (MSE)'), (_lib.LOSS_BCE_SIGMOID, 'CMF term (BCE)')):
    sec = timeit(lambda: ops.train_steps(ut, it, bp['source_user_id'], bp['source_item_id'], None, lab, loss_kind=kind,
                                         reg_weight=0.01, user_dst=gu, item_dst=gi))
    report(f'A3/A16 persistent pointwise steps {nm}, {K} x B=8192', sec, bytes_=K * B * 1044, units=K * B, unit_name='inter')
sec = timeit(lambda: ops.train_steps(ut, it, bp['source_user_id'], bp['source_item_id'], None, lab, loss_kind=_lib.LOSS_BCE_SIGMOID,
                                     reg_weight=0.0, user_dst=gu, item_dst=gi))
report('A16 persistent CMF term, reg_weight = 0 (CMF.yaml default)', sec, bytes_=K * B * 1044, units=K * B, unit_name='inter')

# ---- A4: EMCDR map step (gather -> 64x128 tanh -> 128x64 -> MSE vs gathered target -> backward), b = 8192 -----------
from fake_data import base_config
from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
cfg = base_config(device=dev, latent_factor_model='BPR', source_embedding_size=64, target_embedding_size=64, reg_weight=0.01,
                  mapping_function='non_linear', mlp_hidden_size=[128])
with torch.device(dev):
    m = EMCDR(cfg, ds)
m.set_phase('OVERLAP')
ops.set_table_grad_mode('inplace')
ov = torch.randint(0, ds.num_overlap_user, (8192, 1), device=dev, generator=g)
def map_step():
    loss = m.calculate_loss(Interaction({'overlap': ov}))
    loss.backward()
report('A4 EMCDR map step fwd+bwd (fused MLP kernels, eager launches), b = 8192', timeit(map_step), bytes_=8192 * 1032, flops=8192 * 98304, units=8192, unit_name='inter')
from recbole_cdr_b200.trainer import GraphedTrainStep
gs = GraphedTrainStep(m, Interaction({'overlap': ov}))
report('A4 EMCDR map step fwd+bwd, CUDA-graph replay, b = 8192', timeit(lambda: gs(Interaction({'overlap': ov})), inner=10), bytes_=8192 * 1032, flops=8192 * 98304, units=8192, unit_name='inter')
del m, gs
with torch.device(dev):
    m = EMCDR(dict(cfg, xdr_fused_mlp=True), ds)
m.set_phase('OVERLAP')
report('A4 EMCDR map step fwd+bwd, ONE fused kernel (xdr_fused_mlp), eager, b = 8192', timeit(map_step), bytes_=8192 * 1032, flops=8192 * 98304, units=8192, unit_name='inter')
gs = GraphedTrainStep(m, Interaction({'overlap': ov}))
report('A4 EMCDR map step fwd+bwd, ONE fused kernel, CUDA-graph replay, b = 8192', timeit(lambda: gs(Interaction({'overlap': ov})), inner=10), bytes_=8192 * 1032, flops=8192 * 98304, units=8192, unit_name='inter')
del m, gs

# ---- A14/A15: DTCDR NeuMF BOTH step, D = 64, B = 8192 per domain ------------------------------------------------------
from recbole_cdr_b200.model.cross_domain_recommender.dtcdr import DTCDR
from recbole_cdr_b200.data.synthetic import SyntheticCrossDomainDataset
dsb = SyntheticCrossDomainDataset(IdSpace(500_001, 500_000, 500_000), IdSpace(200_001, 400_000, 400_000))
with torch.device(dev):
    m = DTCDR(base_config(device=dev, embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF', alpha=0.5), dsb)
def both_batch(dsx, Bx, seed):
    b = synthetic.make_batch(dsx, 'source', Bx, seed, dev, pairwise=False)
    b.update(synthetic.make_batch(dsx, 'target', Bx, seed + 1, dev, pairwise=False))
    return Interaction(b)
ib = both_batch(dsb, 8192, 5)
def dt_step():
    m.calculate_loss(ib).backward()
report('A14-15 DTCDR NeuMF BOTH step fwd+bwd (fused MLP kernels, eager launches), 2 x B=8192', timeit(dt_step), bytes_=2 * 8192 * 2068, flops=2 * 8192 * 27700, units=2 * 8192, unit_name='inter')
gs = GraphedTrainStep(m, ib)
report('A14-15 DTCDR NeuMF BOTH step fwd+bwd, CUDA-graph replay', timeit(lambda: gs(ib), inner=10), bytes_=2 * 8192 * 2068, flops=2 * 8192 * 27700, units=2 * 8192, unit_name='inter')
del m, gs
with torch.device(dev):
    m = DTCDR(base_config(device=dev, embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF', alpha=0.5, xdr_fused_mlp=True), dsb)
report('A14-15 DTCDR NeuMF BOTH step fwd+bwd, fused kernel per domain (xdr_fused_mlp), eager', timeit(dt_step), bytes_=2 * 8192 * 2068, flops=2 * 8192 * 27700, units=2 * 8192, unit_name='inter')
gs = GraphedTrainStep(m, ib)
report('A14-15 DTCDR NeuMF BOTH step fwd+bwd, fused kernel per domain, CUDA-graph replay', timeit(lambda: gs(ib), inner=10), bytes_=2 * 8192 * 2068, flops=2 * 8192 * 27700, units=2 * 8192, unit_name='inter')
del m, gs

# ---- A7/A8: CoNet BOTH step at config #3 (5M users / 2M items per domain, 50% user overlap, D = 128, B = 16384) -------
from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
dsc = SyntheticCrossDomainDataset(IdSpace(2_500_001, 2_500_000, 2_500_000), IdSpace(1, 2_000_000, 2_000_000))
with torch.device(dev):
    m = CoNet(base_config(device=dev, embedding_size=128, reg_weight=0.01, mlp_hidden_size=[64, 32, 16, 8]), dsc)
ic = both_batch(dsc, 16384, 9)
def co_step():
    m.calculate_loss(ic).backward()
report('A7-8 CoNet BOTH step fwd+bwd (eager launches), config #3 shape, 2 x B=16384', timeit(co_step, reps=3), bytes_=2 * 16384 * 4116, flops=2 * 16384 * 458000, units=2 * 16384, unit_name='inter')
gs = GraphedTrainStep(m, ic)
report('A7-8 CoNet BOTH step fwd+bwd, CUDA-graph replay', timeit(lambda: gs(ic), reps=3, inner=5), bytes_=2 * 16384 * 4116, flops=2 * 16384 * 458000, units=2 * 16384, unit_name='inter')
del m, gs
torch.cuda.empty_cache()

# ---- A16 at BASELINE config #1 sizes (CMF, bundled ml-1m -> ml-100k after filtering: 6982 users, 3351 items, 1154 overlapped
#      items, no user overlap; B = 2048 per domain): K = 64 BOTH steps = 2 persistent launches, vs the oracle port on the host
from oracle import cdr_oracle as O
dsm = SyntheticCrossDomainDataset(IdSpace(1, 943, 6039), IdSpace(1155, 528, 1668))
cu = torch.randn(dsm.num_total_user, 64, device=dev) * 0.02
ci = torch.randn(dsm.num_total_item, 64, device=dev) * 0.02
gcu, gci = torch.zeros_like(cu), torch.zeros_like(ci)
Kc, Bc = 64, 2048
def cmf_blocks(domain, seed):
    bs = [synthetic.make_batch(dsm, domain, Bc, seed + s, 'cpu', pairwise=False) for s in range(Kc)]
    return {k: torch.stack([b[k] for b in bs]).to(dev) for k in bs[0]}
sb, tb = cmf_blocks('source', 100), cmf_blocks('target', 200)
def cmf_both():
    ops.train_steps(cu, ci, sb['source_user_id'], sb['source_item_id'], None, sb['source_label'].contiguous(),
                    loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=0.0, user_dst=gcu, item_dst=gci, scale=0.5)
    ops.train_steps(cu, ci, tb['target_user_id'], tb['target_item_id'], None, tb['target_label'].contiguous(),
                    loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=0.0, user_dst=gcu, item_dst=gci, scale=0.5)
sec = timeit(cmf_both)
report(f'A16 CMF BOTH steps at config #1 sizes (6982 x 3351 tables, 2 x B=2048), {Kc} steps = 2 launches', sec,
       bytes_=Kc * 2 * Bc * 1044, units=Kc * 2 * Bc, unit_name='inter')
torch.set_num_threads(os.cpu_count())
cuc, cic = torch.nn.Parameter(cu.cpu()), torch.nn.Parameter(ci.cpu())
hb = [{k: v[s].cpu() for k, v in {**sb, **tb}.items()} for s in range(6)]
def cpu_step(b):
    cuc.grad = cic.grad = None
    O.cmf_loss(cuc, cic, b['source_user_id'], b['source_item_id'], b['source_label'], b['target_user_id'], b['target_item_id'],
               b['target_label'], 0.5, 0.0, 0.0).sum().backward()
cpu_step(hb[0])
t0 = time.perf_counter()
for b in hb[1:]:
    cpu_step(b)
cpu_sec = (time.perf_counter() - t0) / 5
line = {'kernel': f'A16 CMF BOTH step, oracle port on {os.cpu_count()} host threads (config #1 sizes)', 'us': cpu_sec * 1e6,
        'Minter_per_s': 2 * Bc / cpu_sec / 1e6}
results.append(line); print(json.dumps(line))
if ONLY:
    json.dump(results, open(os.path.join(ROOT, 'gpurun_out', 'bench_rows.json'), 'w'), indent=1)
    sys.exit(0)
# ---- A10-A12: BiTGCF graph layer at 1/4 of config #4 (0.5M users x 0.25M items per domain, 8M edges, D = 64) ----------
from recbole_cdr_b200.graph import GraphProp, NormAdj, TransferNorm
nu, ni, E = 750_000, 375_000, 8_000_000
rng = np.random.RandomState(0)
r = rng.randint(0, nu, E); c = np.minimum(rng.zipf(1.05, E) - 1, ni - 1)
t0 = time.time(); adj = NormAdj(r, c, nu, ni, dev); build_s = time.time() - t0
X = torch.randn(nu + ni, D, device=dev) * 0.1
nn_ = nu + ni
report(f'A10 spmm_csr L.E  (N={nn_}, nnz={adj.nnz}, D=64; host graph build {build_s:.1f} s)', timeit(lambda: adj.spmm(X), reps=3),
       bytes_=adj.nnz * (12 + 256) + nn_ * 256, units=adj.nnz, unit_name='nnz')
Xg = X.clone().requires_grad_(True)
def prop_fb():
    Xg.grad = None
    GraphProp.apply(Xg, adj).backward(X)
report('A10 graph_layer fwd+bwd (2 SpMM + 3 element-wise)', timeit(prop_fb, reps=3), bytes_=2 * (adj.nnz * (12 + 256) + 2 * nn_ * 256) + 6 * nn_ * 256, units=2 * adj.nnz, unit_name='nnz')
deg = torch.rand(nn_, device=dev) * 10
Y = torch.randn_like(X)
report('A11-12 transfer + normalize fwd (both domains)', timeit(lambda: TransferNorm.apply(X, Y, deg, deg, nu, ni, nu // 2, 1, 0.8, 0.8)), bytes_=6 * nn_ * 256 + 2 * nn_ * 4, units=nn_)

# BiTGCF whole step (2 layers, both domains, concat) on the same graph, graph-captured
# ---- A18: negative draw, 1M users, 16 used items each, 8192 keys x 1 ---------------------------------------------------
from recbole_cdr_b200.sampler import TargetDomainSampler
uu = np.repeat(np.arange(1, 1_000_000), 16); ii = rng.randint(1, 1_000_000, uu.size)
smp = TargetDomainSampler(1_000_000, 1_000_000, uu, ii, device=dev)
keys = torch.randint(1, 1_000_000, (8192,), device=dev, generator=g)
report('A18 neg_sample_uniform 8192 keys (1M items, 16 used/user)', timeit(lambda: smp.sample_by_key_ids(keys, 1, check=False)), units=8192, unit_name='draws')
keys2 = torch.randint(1, 1_000_000, (1 << 20,), device=dev, generator=g)
report('A18 neg_sample_uniform 1M keys', timeit(lambda: smp.sample_by_key_ids(keys2, 1, check=False)), units=1 << 20, unit_name='draws')

json.dump(results, open(os.path.join(ROOT, 'gpurun_out', 'bench_rows.json'), 'w'), indent=1)
