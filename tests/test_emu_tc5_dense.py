"""The tcgen05 dense-layer engine (tc5_dense.cu: forward, input gradient, weight gradient as bf16x3 products with accumulators
in tensor memory) through the CPU CTA emulator, against fp64 torch: every entry point of the dense family, with and without
the cross-stitch operand and the overlap mask, at row counts around the 128-row tile and for every K chunking.  Logic only;
the hardware twin is tests/test_gpu_tc5_dense.py."""
import pytest
import torch

import emu_util


class _engine:
    """xdr_set_dense_engine(e) of the emulator library for the duration of a with block."""

    def __init__(self, e):
        self.e = e

    def __enter__(self):
        import ctypes
        self.L = emu_util.lib()
        self.L.xdr_set_dense_engine.argtypes = [ctypes.c_int]
        self.prev = self.L.xdr_set_dense_engine(self.e)

    def __exit__(self, *exc):
        self.L.xdr_set_dense_engine(self.prev)


def ref_dense(X, W, b, X2, W2, mask, act):
    z = X.double() @ W.double().t()
    if b is not None:
        z = z + b.double()
    if X2 is not None:
        z = z + mask.double().unsqueeze(1) * (X2.double() @ W2.double().t())
    return {0: z, 1: torch.relu(z), 2: torch.tanh(z), 3: torch.sigmoid(z)}[act]


@pytest.mark.parametrize('M,N,K,act,cross', [(128, 16, 16, 1, False), (300, 64, 256, 1, True), (257, 32, 64, 2, False),
                                             (129, 128, 64, 0, True), (640, 48, 48, 3, False), (200, 16, 32, 1, True),
                                             (384, 64, 128, 1, True), (130, 32, 192, 2, False)])
def test_dense_layer_on_tcgen05_matches_fp64(M, N, K, act, cross):
    from recbole_cdr_b200 import _lib
    g = torch.Generator().manual_seed(M + N + K)
    X, W = torch.randn(M, K, generator=g) * 0.5, torch.randn(N, K, generator=g) * 0.2
    b = torch.randn(N, generator=g) * 0.1
    X2 = torch.randn(M, K, generator=g) * 0.5 if cross else None
    W2 = torch.randn(N, K, generator=g) * 0.2 if cross else None
    ids = torch.randint(0, 100, (M,), generator=g) if cross else None
    mask = (ids < 40) if cross else None
    dY = torch.randn(M, N, generator=g)
    leaves = [t.clone().double().requires_grad_(True) for t in (X, W, b)] + \
        ([X2.clone().double().requires_grad_(True), W2.clone().double().requires_grad_(True)] if cross else [None, None])
    want = ref_dense(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], mask, act)
    want.backward(dY.double())
    with emu_util.patched_ops(sms=2, seed=1) as ops, _engine(1):
        c = [t.clone().requires_grad_(True) for t in (X, W, b)] + \
            ([X2.clone().requires_grad_(True), W2.clone().requires_grad_(True)] if cross else [None, None])
        Y = ops.dense(c[0], c[1], c[2], act, c[3], c[4], ids, 40)
        Y.backward(dY)
    scale = lambda t: max(1e-6, float(t.abs().max()))
    torch.testing.assert_close(Y.detach().double(), want.detach(), rtol=1e-4, atol=2e-5 * scale(want))
    names = ('X', 'W', 'b', 'X2', 'W2')
    for nm, got, ref in zip(names, c, leaves):
        if got is None:
            continue
        torch.testing.assert_close(got.grad.double(), ref.grad, rtol=2e-4, atol=5e-5 * scale(ref.grad), msg=lambda s: f'{nm}: {s}')


def test_dense_engine_switch_and_fallback_shapes():
    """xdr_set_dense_engine(1) moves a qualifying shape from the fp32 FMA kernels to tcgen05 (results agree); shapes the
    engine does not take (N = 8, K = 20, M < 128) never reach it."""
    g = torch.Generator().manual_seed(3)
    X, W = torch.randn(256, 64, generator=g), torch.randn(32, 64, generator=g) * 0.2
    with emu_util.patched_ops(sms=2) as ops:
        y0 = ops.dense(X, W, None, 0)       # default engine: fp32 FMA
        with _engine(1):
            y1 = ops.dense(X, W, None, 0)
            for (m, n, k) in ((64, 32, 64), (256, 8, 64), (256, 32, 20)):   # shapes the engine does not take
                x, w = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g)
                torch.testing.assert_close(ops.dense(x, w, None, 0), x @ w.t(), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(y1, y0, rtol=1e-4, atol=1e-4)
        assert not torch.equal(y1, y0)      # different arithmetic (bf16x6 on tensor cores vs fp32 FMA): the switch switches


def test_engine_2_takes_only_the_wide_forward_and_input_gradient():
    """xdr_set_dense_engine(2): tcgen05 where it measured faster on a B200 (K >= 192 and N >= 64: forward and input gradient),
    the fp32 tiles for narrow layers and for every weight gradient -- told apart by whose arithmetic the result carries."""
    g = torch.Generator().manual_seed(5)
    dY = {}
    with emu_util.patched_ops(sms=2) as ops:
        def run(e, K, N):
            X, W = torch.randn(256, K, generator=torch.Generator().manual_seed(K + N)), \
                torch.randn(N, K, generator=torch.Generator().manual_seed(K * N)) * 0.2
            X.requires_grad_(True)
            W.requires_grad_(True)
            d = dY.setdefault((K, N), torch.randn(256, N, generator=g))
            with _engine(e):
                Y = ops.dense(X, W, None, 0)
                Y.backward(d)
            return Y.detach(), X.grad, W.grad
        wide = {e: run(e, 256, 64) for e in (0, 1, 2)}
        assert torch.equal(wide[2][0], wide[1][0]) and not torch.equal(wide[2][0], wide[0][0])   # forward: tcgen05
        assert torch.equal(wide[2][1], wide[1][1]) and not torch.equal(wide[2][1], wide[0][1])   # input gradient: tcgen05
        torch.testing.assert_close(wide[2][2], wide[0][2], rtol=1e-5, atol=1e-5)                  # weight gradient: fp32 tiles
        assert not torch.equal(wide[1][2], wide[0][2])
        narrow = {e: run(e, 64, 32) for e in (0, 1, 2)}
        assert torch.equal(narrow[2][0], narrow[0][0]) and torch.equal(narrow[2][1], narrow[0][1])   # narrow layer: fp32 tiles
        assert not torch.equal(narrow[1][0], narrow[0][0])


@pytest.mark.parametrize('engine', [0, 1])
@pytest.mark.parametrize('M,N,K', [(200, 32, 64), (129, 16, 32)])
def test_cross_pair_node_equals_two_dense_nodes(engine, M, N, K):
    """ops.cross_pair (both directions of a cross-stitch layer as one autograd node: the second input-gradient product
    ACCUMULATES into the first one's result, both dH products add into one destination) against two ops.dense nodes whose
    gradient halves autograd adds -- on both engines (the accumulating input-gradient form of each)."""
    g = torch.Generator().manual_seed(M + N + K + engine)
    mk = lambda *s, sc=1.0: torch.randn(*s, generator=g) * sc
    base = [mk(M, K, sc=0.5), mk(M, K, sc=0.5), mk(N, K, sc=0.2), mk(N, sc=0.1), mk(N, K, sc=0.2), mk(N, sc=0.1), mk(N, K, sc=0.2)]
    ids = torch.randint(0, 100, (M,), generator=g)
    d_s, d_t = mk(M, N), mk(M, N)
    res = []
    with emu_util.patched_ops(sms=2) as ops, _engine(engine):
        for pair in (True, False):
            x_s, x_t, Ws, bs, Wt, bt, H = [t.clone().requires_grad_(True) for t in base]
            if pair:
                h_s, h_t = ops.cross_pair(x_s, x_t, Ws, bs, Wt, bt, H, ids, 40, 1)
            else:
                h_s = ops.dense(x_s, Ws, bs, 1, x_t, H, ids, 40)
                h_t = ops.dense(x_t, Wt, bt, 1, x_s, H, ids, 40)
            torch.autograd.backward([h_s, h_t], [d_s, d_t])
            res.append([h_s.detach(), h_t.detach()] + [t.grad for t in (x_s, x_t, Ws, bs, Wt, bt, H)])
        # one output unused: its half of every gradient is absent, not garbage
        x_s, x_t, Ws, bs, Wt, bt, H = [t.clone().requires_grad_(True) for t in base]
        h_s, _ = ops.cross_pair(x_s, x_t, Ws, bs, Wt, bt, H, ids, 40, 1)
        h_s.backward(d_s)
        y_s, y_t, Vs, cs, Vt, ct, G = [t.clone().requires_grad_(True) for t in base]
        ops.dense(y_s, Vs, cs, 1, y_t, G, ids, 40).backward(d_s)
        for a, b in ((x_s, y_s), (x_t, y_t), (Ws, Vs), (bs, cs), (H, G)):
            torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-6)
        assert Wt.grad is None or not bool(Wt.grad.any())
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    for a, b in zip(res[0][2:], res[1][2:]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)


def test_per_call_engine_reaches_the_backward_and_does_not_leak():
    """``ops.dense(..., engine=e)`` / ``ops.cross_pair(..., engine=e)`` (how CoNet picks tcgen05 for its own layers): the forward
    AND the backward launches -- which autograd runs later, outside the caller's code -- take engine ``e``; the library's own
    setting is what it was before and after."""
    import ctypes
    g = torch.Generator().manual_seed(11)
    X0, W0 = torch.randn(256, 64, generator=g) * 0.5, torch.randn(32, 64, generator=g) * 0.2
    d = torch.randn(256, 32, generator=g)
    L = emu_util.lib()
    L.xdr_set_dense_engine.argtypes = [ctypes.c_int]

    def run(ops, global_engine, call_engine):
        X, W = X0.clone().requires_grad_(True), W0.clone().requires_grad_(True)
        with _engine(global_engine):
            Y = ops.dense(X, W, None, 1, engine=call_engine)
            assert L.xdr_set_dense_engine(global_engine) == global_engine     # restored right after the forward launch
            Y.backward(d)
            assert L.xdr_set_dense_engine(global_engine) == global_engine
        return Y.detach(), X.grad, W.grad

    with emu_util.patched_ops(sms=2) as ops:
        fma, tc5 = run(ops, 0, None), run(ops, 1, None)
        assert not torch.equal(fma[0], tc5[0]) and not torch.equal(fma[2], tc5[2])
        for a, b in zip(run(ops, 0, 1), tc5):      # per-call tcgen05 under a library default of fp32 tiles
            assert torch.equal(a, b)
        for a, b in zip(run(ops, 1, 0), fma):      # and the other way round
            assert torch.equal(a, b)
