"""Per-shape timing of the dense-layer entry points on both engines (fp32 FMA vs tcgen05 bf16x6), CUDA events, warm, 20 calls
back to back.  Shapes: the layers of CoNet config #3, the EMCDR map MLP, the NeuMF towers.  One JSON line per shape."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'recbole-cdr_b200')); sys.path.insert(0, ROOT)
import torch
from recbole_cdr_b200 import _lib
from recbole_cdr_b200._lib import call, ptr, cur_stream

dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
SHAPES = [('conet L0 (cross)', 16384, 64, 256, True), ('conet L1 (cross)', 16384, 32, 64, True), ('conet L2 (cross)', 16384, 16, 32, True),
          ('map L1', 8192, 128, 64, False), ('map L2', 8192, 64, 128, False), ('neumf L1', 8192, 32, 128, False),
          ('neumf L2', 8192, 16, 32, False), ('map L1 b=32768', 32768, 128, 64, False)]


def timeit(fn, inner=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best * 1e3


for name, M, N, K, cross in SHAPES:
    X, X2 = torch.randn(M, K, device=dev, generator=g), torch.randn(M, K, device=dev, generator=g)
    W, W2 = torch.randn(N, K, device=dev, generator=g) * 0.1, torch.randn(N, K, device=dev, generator=g) * 0.1
    b = torch.randn(N, device=dev, generator=g)
    ids = torch.randint(0, 100, (M,), device=dev, generator=g)
    Y, dZ, dX = torch.empty(M, N, device=dev), torch.randn(M, N, device=dev, generator=g), torch.empty(M, K, device=dev)
    dW, db = torch.zeros(N, K, device=dev), torch.zeros(N, device=dev)
    s = cur_stream()
    row = {'layer': name, 'M': M, 'N': N, 'K': K}
    for eng in (0, 1):
        _lib._lib.xdr_set_dense_engine(eng)
        t = 'tc5' if eng else 'fma'
        row[f'fwd_{t}_us'] = round(timeit(lambda: call('xdr_dense_fwd', ptr(X), ptr(W), ptr(b), ptr(X2) if cross else None,
                                                       ptr(W2) if cross else None, ptr(ids) if cross else None, 40, 1, ptr(Y), M, N, K, s)), 2)
        row[f'bwd_input_{t}_us'] = round(timeit(lambda: call('xdr_dense_bwd_input', ptr(dZ), ptr(W), None, 0, ptr(dX), M, N, K, 0, s)), 2)
        row[f'bwd_weight_{t}_us'] = round(timeit(lambda: call('xdr_dense_bwd_weight', ptr(dZ), ptr(X), None, 0, ptr(dW), ptr(db), M, N, K, s)), 2)
    _lib._lib.xdr_set_dense_engine(1)
    print(json.dumps(row), flush=True)
