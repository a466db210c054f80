"""Debug: timeline of the persistent multi-step kernel from in-kernel globaltimer stamps (see xdr_debug_set_steps_trace)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'recbole-cdr_b200')); sys.path.insert(0, ROOT)
import torch
from recbole_cdr_b200 import _lib, ops
from recbole_cdr_b200.data import synthetic

fused = len(sys.argv) > 1 and sys.argv[1] == 'fused'
lazy = len(sys.argv) > 1 and sys.argv[1] == 'lazy'
dev = torch.device('cuda', 0)
ds = synthetic.emcdr_scale(1_000_000)
K, B, D = 64, 8192, 64
torch.manual_seed(0)
ut = torch.randn(ds.num_total_user, D, device=dev) * 0.01
it = torch.randn(ds.num_total_item, D, device=dev) * 0.01
ids = torch.stack([torch.stack([synthetic.make_batch(ds, 'source', B, 1 + s, 'cpu')[k] for k in
                                ('source_user_id', 'source_item_id', 'neg_source_item_id')]) for s in range(K)]).to(dev)
gu, gi = (ut, it) if fused else (torch.zeros_like(ut), torch.zeros_like(it))
G = 147
trace = torch.zeros(K * G * 8, dtype=torch.int64, device=dev)
lib = ctypes.CDLL(_lib.LIB_PATH)
lib.xdr_debug_set_steps_trace.argtypes = [ctypes.c_void_p]
for rep in range(2):
    lib.xdr_debug_set_steps_trace(trace.data_ptr() if rep == 1 else None)
    kw = dict(touch=ops.TouchMap(ut.shape[0], it.shape[0], dev), fresh=True) if lazy else {}
    ops.train_steps(ut, it, ids[:, 0], ids[:, 1], ids[:, 2], reg_weight=0.01, user_dst=gu, item_dst=gi,
                    scale=-0.01 if fused else 1.0, **kw)
    torch.cuda.synchronize()
lib.xdr_debug_set_steps_trace(None)
t = trace.view(K, G, 8).cpu().double()
t0 = t[:, :, 2][t[:, :, 2] > 0].min()
t = (t - t0) / 1e3  # us
names = ['cta partial ready', 'all partials seen', 'task0 rows requested', 'task0 scored', 'task0 waits norms',
         'norms arrived', 'scatter issued']
print('mode:', 'fused sgd (dst = tables)' if fused else ('lazily zeroed grad tables' if lazy else 'grad tables'))
print('step | ' + ' | '.join(f'{n[:18]:>18s}' for n in names) + '   (median over CTAs; max for col 0/1), us since start')
for s in list(range(0, 12)) + list(range(40, 46)):
    row = []
    for k in range(7):
        col = t[s, :, k]
        col = col[col > -1e6]
        row.append(f'{col.median():8.2f}/{col.max():8.2f}' if col.numel() else '-')
    print(f'{s:4d} | ' + ' | '.join(f'{x:>18s}' for x in row))
d = t[1:, :, 1].max(dim=1).values - t[:-1, :, 1].max(dim=1).values
print('median step period (all partials seen, max over CTAs): %.2f us' % d.median())
print('exchange latency  = all partials seen - last CTA partial ready: %.2f us (median over steps)' %
      (t[:, :, 1].max(dim=1).values - t[:, :, 0].max(dim=1).values).median())
print('skew of partial-ready across CTAs (max - min): %.2f us' % (t[:, :, 0].max(dim=1).values - t[:, :, 0].min(dim=1).values).median())
print('rows requested -> scored (task 0): %.2f us median, %.2f p95' % ((t[:, :, 3] - t[:, :, 2]).median(), (t[:, :, 3] - t[:, :, 2]).flatten().quantile(0.95)))
print('scored -> norms arrived (task 0): %.2f us median' % (t[:, :, 5] - t[:, :, 3]).median())
print('norms arrived -> scatter issued (scatter warp 0): %.2f us median' % (t[:, :, 6] - t[:, :, 5]).median())
# plain timing of both destination modes
if lazy:
    sys.exit(0)
for mode, (du, di, sc) in {'grad tables': (torch.zeros_like(ut), torch.zeros_like(it), 1.0), 'fused sgd': (ut, it, -1e-3)}.items():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(4):
        e0.record()
        ops.train_steps(ut, it, ids[:, 0], ids[:, 1], ids[:, 2], reg_weight=0.01, user_dst=du, item_dst=di, scale=sc)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f'{mode:12s}: {best * 1e3 / K:6.2f} us/step  ({B * K / best / 1e6:.2f} G interactions/s)')
