"""Diagnostic (GPU): the fused CoNet tower kernel on config #3's shape with Zipf user ids, against an fp64 oracle -- where do
the largest errors sit (which rows, which 64-column chunk, how many duplicates), and how far is the fp32 oracle itself from
fp64?  Prints one JSON object."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'recbole-cdr_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch
from oracle import cdr_oracle as O
from recbole_cdr_b200 import ops
from test_gpu_kernels import rand_ids, rand_table

dev = torch.device('cuda', 0)
batch, dim, hidden, want = int(os.environ.get('B', 16384)), 128, [64, 32, 16, 8], 1
gen = torch.Generator().manual_seed(5)
n_u, n_i, n_ov = 5000, 3000, 2500
names = ('source_user', 'source_item', 'target_user', 'target_item')
tabs = {k: rand_table(n_u if 'user' in k else n_i, dim, 300 + j, 0.3) for j, k in enumerate(names)}
dims = [2 * dim] + hidden
mk = lambda a, b, s=0.25: torch.randn(b, a, generator=gen) * s
P = dict(ws=[mk(a, b) for a, b in zip(dims[:-1], dims[1:])], wt=[mk(a, b) for a, b in zip(dims[:-1], dims[1:])],
         h=[mk(a, b) for a, b in zip(dims[:-1], dims[1:])],
         bs=[torch.randn(b, generator=gen) * 0.1 for b in dims[1:]], bt=[torch.randn(b, generator=gen) * 0.1 for b in dims[1:]],
         out_s_w=mk(dims[-1], 1, 0.5), out_s_b=torch.randn(1, generator=gen) * 0.1,
         out_t_w=mk(dims[-1], 1, 0.5), out_t_b=torch.randn(1, generator=gen) * 0.1)
user, item = rand_ids(batch, n_u, 7, 1.3), rand_ids(batch, n_i, 8)
label = (torch.rand(batch, generator=gen) < 0.5).float()


def oracle(dtype):
    lt = {k: v.clone().to(dtype).requires_grad_(True) for k, v in tabs.items()}
    lp = {k: ([x.clone().to(dtype).requires_grad_(True) for x in v] if isinstance(v, list) else v.clone().to(dtype).requires_grad_(True))
          for k, v in P.items()}
    ps, pt = O.conet_towers(lt, user, item, lp, True, n_ov)
    ref = O.bce_loss(ps if want == 0 else pt, label.to(dtype))
    ref.backward()
    return lt, lp, ref


lt64, lp64, ref64 = oracle(torch.float64)
lt32, lp32, ref32 = oracle(torch.float32)
out = {'batch': batch, 'runs': []}
for rep in range(int(os.environ.get('REPS', 3))):
    ct = {k: v.to(dev).requires_grad_(True) for k, v in tabs.items()}
    cp = {k: ([x.to(dev).requires_grad_(True) for x in v] if isinstance(v, list) else v.to(dev).requires_grad_(True)) for k, v in P.items()}
    loss = ops.conet_tower_loss(want, False, n_ov, user.to(dev), item.to(dev), label.to(dev), tuple(ct[k] for k in names),
                                cp['out_t_w'], cp['out_t_b'], cp['ws'], cp['bs'], cp['wt'], cp['bt'], cp['h'])
    loss.backward()
    torch.cuda.synchronize()
    run = {'loss': float(loss), 'loss64': float(ref64)}
    cnt = np.bincount(user.numpy(), minlength=n_u)
    for name in names:
        t = lt64[name].grad.numpy()
        g = ct[name].grad.cpu().numpy().astype(np.float64)
        o = lt32[name].grad.numpy().astype(np.float64)
        eg, eo = np.abs(g - t), np.abs(o - t)
        rows = np.argsort(-eg.max(1))[:4]
        run[name] = {'max_grad': float(np.abs(t).max()), 'kernel_err': float(eg.max()), 'fp32_oracle_err': float(eo.max()),
                     'worst_rows': [{'row': int(r), 'dups': int(cnt[r]) if 'user' in name else None, 'err': float(eg[r].max()),
                                     'row_mag': float(np.abs(t[r]).max()),
                                     'err_by_chunk': [float(eg[r, c * 64:(c + 1) * 64].max()) for c in range(dim // 64)],
                                     'n_bad_cols': int((eg[r] > 1e-4 * np.abs(t).max()).sum())} for r in rows]}
    for key in ('ws', 'wt', 'h'):
        t = lp64[key][0].grad.numpy(); g = cp[key][0].grad.cpu().numpy().astype(np.float64)
        run[f'd{key}0_err_rel'] = float(np.abs(g - t).max() / np.abs(t).max())
    out['runs'].append(run)
print(json.dumps(out))
