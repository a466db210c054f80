"""GPU: the tcgen05 dense-layer engine (tc5_dense.cu) against fp64 torch on the device -- forward, input gradient and weight
gradient of `act(X W^T + b + m (X2 W2^T))` at the BASELINE shapes (CoNet config #3's layer 0: 16384 x 256 -> 64 with the
cross-stitch operand; the EMCDR map MLP; NeuMF towers) and at ragged row counts.  CPU twin: tests/test_emu_tc5_dense.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda', 0)


@pytest.fixture(autouse=True)
def tcgen05_engine():
    """The engine is opt-in (xdr_set_dense_engine): switch it on for the tests of this file."""
    from recbole_cdr_b200 import _lib
    prev = _lib._lib.xdr_set_dense_engine(1)
    yield
    _lib._lib.xdr_set_dense_engine(prev)


@pytest.mark.parametrize('M,N,K,act,cross', [(128, 16, 16, 1, False), (16384, 64, 256, 1, True), (8192, 128, 64, 2, False),
                                             (8192, 64, 128, 0, False), (16384, 32, 64, 1, True), (16384, 16, 32, 1, True),
                                             (1000, 48, 48, 3, False), (4097, 32, 192, 2, True), (8192, 32, 128, 1, False)])
def test_dense_layer_on_tcgen05_matches_fp64(M, N, K, act, cross):
    from recbole_cdr_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev())
    X, W, b = mk(M, K, sc=0.5), mk(N, K, sc=0.2), mk(N, sc=0.1)
    X2, W2 = (mk(M, K, sc=0.5), mk(N, K, sc=0.2)) if cross else (None, None)
    ids = torch.randint(0, 100, (M,), generator=g).to(dev()) if cross else None
    dY = mk(M, N)
    def pre(xs):
        z_ = xs[0] @ xs[1].t() + xs[2]
        return z_ + (ids < 40).double().unsqueeze(1) * (xs[3] @ xs[4].t()) if cross else z_
    if act == 1:   # ReLU is not differentiable at 0: rows with a pre-activation within rounding distance of it are rescaled
        for _ in range(6):
            edge = (pre([t.double() if t is not None else None for t in (X, W, b, X2, W2)]).abs() < 1e-4).any(1)
            if not bool(edge.any()):
                break
            X[edge] *= 1.37
    ref = [t.double().requires_grad_(True) if t is not None else None for t in (X, W, b, X2, W2)]
    z = pre(ref)
    want = {0: z, 1: torch.relu(z), 2: torch.tanh(z), 3: torch.sigmoid(z)}[act]
    want.backward(dY.double())
    c = [t.clone().requires_grad_(True) if t is not None else None for t in (X, W, b, X2, W2)]
    Y = ops.dense(c[0], c[1], c[2], act, c[3], c[4], ids, 40)
    Y.backward(dY)
    torch.cuda.synchronize()
    scale = lambda t: max(1e-6, float(t.detach().abs().max()))
    torch.testing.assert_close(Y.detach().double(), want.detach(), rtol=1e-4, atol=2e-5 * scale(want))
    for nm, got, r in zip(('X', 'W', 'b', 'X2', 'W2'), c, ref):
        if got is None:
            continue
        torch.testing.assert_close(got.grad.double(), r.grad, rtol=2e-4, atol=5e-5 * scale(r.grad), msg=lambda s: f'{nm}: {s}')


def test_engine_switch_changes_the_arithmetic_not_the_result():
    from recbole_cdr_b200 import _lib, ops
    g = torch.Generator().manual_seed(3)
    X, W = torch.randn(4096, 64, generator=g).to(dev()), (torch.randn(32, 64, generator=g) * 0.2).to(dev())
    y1 = ops.dense(X, W, None, 0)
    prev = _lib._lib.xdr_set_dense_engine(0)
    try:
        y0 = ops.dense(X, W, None, 0)
    finally:
        _lib._lib.xdr_set_dense_engine(prev)
    torch.cuda.synchronize()
    assert prev == 1   # (the fixture switched it on)
    torch.testing.assert_close(y1, y0, rtol=1e-4, atol=1e-4)
    assert not torch.equal(y1, y0)


@pytest.mark.parametrize('streams', [0, 4])
@pytest.mark.parametrize('engine', [0, 1, 2])
@pytest.mark.parametrize('M,N,K', [(32768, 64, 256), (4097, 32, 64), (333, 16, 32)])
def test_cross_pair_node_equals_two_dense_nodes(engine, M, N, K, streams):
    """ops.cross_pair (both directions of a CoNet cross-stitch layer as one autograd node; the second input-gradient product
    accumulates into the first one's result -- ``xdr_dense_bwd_input(accumulate=1)`` -- and both dH products add into one
    destination) against two ops.dense nodes, on every engine setting.  CPU twin: tests/test_emu_tc5_dense.py."""
    from recbole_cdr_b200 import _lib, ops
    g = torch.Generator().manual_seed(M + N + K + engine)
    mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev())
    base = [mk(M, K, sc=0.5), mk(M, K, sc=0.5), mk(N, K, sc=0.2), mk(N, sc=0.1), mk(N, K, sc=0.2), mk(N, sc=0.1), mk(N, K, sc=0.2)]
    ids = torch.randint(0, 100, (M,), generator=g).to(dev())
    d_s, d_t = mk(M, N), mk(M, N)
    res = []
    prev = _lib._lib.xdr_set_dense_engine(engine)
    prev_streams = ops.set_cross_streams(streams)    # (4: the node's independent launches on parallel streams)
    try:
        for pair in (True, False):
            x_s, x_t, Ws, bs, Wt, bt, H = [t.clone().requires_grad_(True) for t in base]
            if pair:
                h_s, h_t = ops.cross_pair(x_s, x_t, Ws, bs, Wt, bt, H, ids, 40, 1)
            else:
                h_s = ops.dense(x_s, Ws, bs, 1, x_t, H, ids, 40)
                h_t = ops.dense(x_t, Wt, bt, 1, x_s, H, ids, 40)
            torch.autograd.backward([h_s, h_t], [d_s, d_t])
            torch.cuda.synchronize()
            res.append([h_s.detach(), h_t.detach()] + [t.grad for t in (x_s, x_t, Ws, bs, Wt, bt, H)])
    finally:
        _lib._lib.xdr_set_dense_engine(prev)
        ops.set_cross_streams(prev_streams)
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    for nm, a, b in zip(('x_s', 'x_t', 'Ws', 'bs', 'Wt', 'bt', 'H'), res[0][2:], res[1][2:]):
        scale = max(1e-6, float(b.abs().max()))   # (weight gradients are sums of atomics: order differs from run to run)
        torch.testing.assert_close(a, b, rtol=1e-4, atol=2e-5 * scale, msg=lambda s: f'd{nm}: {s}')
