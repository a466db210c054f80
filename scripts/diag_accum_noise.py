"""Diagnostic (GPU): how far are the scatter-added gradient tables from an fp64 accumulation of the same contributions, for the
plain and the lazily zeroed destination modes, after one launch and after a second accumulating launch?  Zipf ids (hot rows)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'recbole-cdr_b200')):
    sys.path.insert(0, p)
import numpy as np
import torch
from recbole_cdr_b200 import ops
from test_gpu_lazy_tables import make, dev

nu, ni, dim, K, B, zipf = 5000, 7000, 64, 25, 8192, 1.2
ut, it, u, ip, ineg = make(nu, ni, dim, K, B, 3, zipf)
# fp64 truth of the reg-free BPR gradients (reg_weight 0 keeps the oracle simple: no batch-wide norms)
U, I = ut.double(), it.double()
gu64, gi64 = torch.zeros_like(U), torch.zeros_like(I)
for k in range(K):
    eu, ep, en = U[u[k]], I[ip[k]], I[ineg[k]]
    x = (eu * ep).sum(1) - (eu * en).sum(1)
    s = torch.sigmoid(x)
    c = (-(s * (1 - s)) / (1e-10 + s) / B).unsqueeze(1)
    gu64.index_add_(0, u[k], c * (ep - en))
    gi64.index_add_(0, ip[k], c * eu)
    gi64.index_add_(0, ineg[k], -c * eu)
out = {}
cnt = torch.bincount(u.reshape(-1), minlength=nu)


def err(tag, gu, gi, mult):
    eu_, ei_ = (gu.double() - mult * gu64).abs(), (gi.double() - mult * gi64).abs()
    r = int(eu_.max(1).values.argmax())
    out[tag] = {'user_err': float(eu_.max()), 'user_max': float((mult * gu64).abs().max()), 'worst_user_row': r,
                'worst_row_dups': int(cnt[r]), 'item_err': float(ei_.max()), 'item_max': float((mult * gi64).abs().max()),
                'n_user_elems_over_1e-6_of_max': int((eu_ > 1e-6 * float(gu64.abs().max())).sum())}


for rep in range(2):
    _, gu, gi = ops.train_steps(ut, it, u, ip, ineg, reg_weight=0.0)
    torch.cuda.synchronize(); err(f'plain_1x_rep{rep}', gu, gi, 1)
    ops.train_steps(ut, it, u, ip, ineg, reg_weight=0.0, user_dst=gu, item_dst=gi)
    torch.cuda.synchronize(); err(f'plain_2x_rep{rep}', gu, gi, 2)
    tm = ops.TouchMap(nu, ni, dev())
    gu, gi = torch.full_like(ut, 7.0), torch.full_like(it, -3.0)
    ops.train_steps(ut, it, u, ip, ineg, reg_weight=0.0, user_dst=gu, item_dst=gi, touch=tm, fresh=True)
    torch.cuda.synchronize()
    tu, ti = tm.touched()
    gu[~tu] = 0; gi[~ti] = 0
    err(f'lazy_1x_rep{rep}', gu, gi, 1)
    ops.train_steps(ut, it, u, ip, ineg, reg_weight=0.0, user_dst=gu, item_dst=gi, touch=tm, fresh=False)
    torch.cuda.synchronize(); err(f'lazy_2x_rep{rep}', gu, gi, 2)
print(json.dumps(out, indent=1))
