"""One profiled launch of every hot-path kernel at the BASELINE shapes, for `ncu --set full --profile-from-start off`:
every op runs once un-profiled (warm-up), then once between cudaProfilerStart / cudaProfilerStop.

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/rows python scripts/ncu_rows.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'recbole-cdr_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from recbole_cdr_b200 import _lib, ops
from recbole_cdr_b200.data import Interaction, synthetic
from recbole_cdr_b200.data.idspace import IdSpace
from recbole_cdr_b200.data.synthetic import SyntheticCrossDomainDataset
from fake_data import base_config

dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
rng = np.random.RandomState(0)
jobs = []   # (name, fn)

# A1 gather / scatter-add: 2M x 64 table, 1M random rows
N, D, n = 2_000_001, 64, 1 << 20
tab = torch.randn(N, D, device=dev) * 0.01
idx = torch.randint(0, N, (n,), device=dev, generator=g)
out = torch.empty(n, D, device=dev)
grad = torch.zeros_like(tab)
jobs.append(('gather_rows', lambda: ops.gather_rows_raw(tab, idx, out)))
jobs.append(('scatter_add_rows', lambda: ops.scatter_add_rows_raw(grad, idx, out)))

# A1-A3/A17 the headline: 20 BPR steps in one persistent launch (config #2)
ds = synthetic.emcdr_scale(1_000_000)
ut = torch.randn(ds.num_total_user, D, device=dev) * 0.01
it = torch.randn(ds.num_total_item, D, device=dev) * 0.01
gu, gi = torch.zeros_like(ut), torch.zeros_like(it)
K, B = 20, 8192
bs = [synthetic.make_batch(ds, 'source', B, 1 + s, 'cpu') for s in range(K)]
ids = torch.stack([torch.stack([b[k] for k in ('source_user_id', 'source_item_id', 'neg_source_item_id')]) for b in bs]).to(dev)
jobs.append(('train_steps K=20', lambda: ops.train_steps(ut, it, ids[:, 0], ids[:, 1], ids[:, 2], reg_weight=0.01, user_dst=gu, item_dst=gi)))

# A4 map step (tcgen05 kernel, the default) and its composed form
from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
ops.set_table_grad_mode('inplace')
cfg = base_config(device=dev, latent_factor_model='BPR', source_embedding_size=64, target_embedding_size=64, reg_weight=0.01,
                  mapping_function='non_linear', mlp_hidden_size=[128])
ov = torch.randint(0, ds.num_overlap_user, (8192, 1), device=dev, generator=g)
with torch.device(dev):
    m_map = EMCDR(cfg, ds)
    m_map_c = EMCDR(dict(cfg, xdr_fused_mlp=False), ds)
for mm in (m_map, m_map_c):
    mm.set_phase('OVERLAP')
jobs.append(('EMCDR map step (tc5_mlp)', lambda: m_map.calculate_loss(Interaction({'overlap': ov})).backward()))
jobs.append(('EMCDR map step (composed)', lambda: m_map_c.calculate_loss(Interaction({'overlap': ov})).backward()))

# A7-A8 CoNet BOTH step at config #3
from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
dsc = SyntheticCrossDomainDataset(IdSpace(2_500_001, 2_500_000, 2_500_000), IdSpace(1, 2_000_000, 2_000_000))
with torch.device(dev):
    m_co = CoNet(base_config(device=dev, embedding_size=128, reg_weight=0.01, mlp_hidden_size=[64, 32, 16, 8]), dsc)
def both_batch(dsx, Bx, seed):
    b = synthetic.make_batch(dsx, 'source', Bx, seed, dev, pairwise=False)
    b.update(synthetic.make_batch(dsx, 'target', Bx, seed + 1, dev, pairwise=False))
    return Interaction(b)
ic = both_batch(dsc, 16384, 9)
jobs.append(('CoNet BOTH step', lambda: m_co.calculate_loss(ic).backward()))

# A9-A12 graph kernels: 1M x 0.5M bipartite graph, 8M edges (a quarter of one config #4 domain: ncu replays every kernel ~40 times)
from recbole_cdr_b200.graph import NormAdj, GraphProp, TransferNorm
nu, ni, E = 1_000_000, 500_000, 8_000_000
r = rng.randint(0, nu, E); c = np.minimum(rng.zipf(1.05, E) - 1, ni - 1)
adj = NormAdj(r, c, nu, ni, dev)
X = torch.randn(nu + ni, D, device=dev) * 0.1
Xg = X.clone().requires_grad_(True)
def prop_fb():
    Xg.grad = None
    GraphProp.apply(Xg, adj).backward(X)
jobs.append(('graph layer fwd+bwd', prop_fb))
deg = torch.rand(nu + ni, device=dev) * 10
Y = torch.randn_like(X)
jobs.append(('transfer + normalise', lambda: TransferNorm.apply(X, Y, deg, deg, nu, ni, nu // 2, 1, 0.8, 0.8)))

# A18 negative draw
from recbole_cdr_b200.sampler import TargetDomainSampler
uu = np.repeat(np.arange(1, 1_000_000), 16); ii = rng.randint(1, 1_000_000, uu.size)
smp = TargetDomainSampler(1_000_000, 1_000_000, uu, ii, device=dev)
keys2 = torch.randint(1, 1_000_000, (1 << 20,), device=dev, generator=g)
jobs.append(('neg_sample 1M keys', lambda: smp.sample_by_key_ids(keys2, 1, check=False)))

# F1 row-sparse optimizer, F2 fused top-k
stamp = torch.zeros(N, dtype=torch.int32, device=dev)
s1, s2 = torch.zeros_like(tab), torch.zeros_like(tab)
oid = idx[:3 * 8192].contiguous()
jobs.append(('row-sparse lazy Adam', lambda: ops.sparse_optim_rows(_lib.OPT_LAZY_ADAM, tab, grad, oid, stamp, 1, 1e-3, state1=s1, state2=s2, adam_t=1)))
U = torch.randn(4096, D, device=dev) * 0.1
jobs.append(('full_sort_topk tc5', lambda: ops.full_sort_topk(U, it, 20, engine='tc5')))

ONLY = os.environ.get('XDR_NCU_JOBS')   # comma-separated substrings of job names (default: all)
if ONLY:
    jobs = [(n_, f_) for n_, f_ in jobs if any(k_ in n_ for k_ in ONLY.split(','))]
for name, fn in jobs:      # warm-up, un-profiled
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for name, fn in jobs:
    fn()
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('profiled', [n for n, _ in jobs])
