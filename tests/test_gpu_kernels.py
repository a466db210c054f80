"""GPU parity of the libxdr kernels against the CPU oracle on seeded inputs (run on the B200: -m gpu).

Bar (BASELINE.json north_star): gathered rows / index side BIT-EXACT; fp32 loss within 1e-4 relative; per-row
gradients within 1e-4 relative of the oracle's dense autograd gradients (duplicates summed, order-free)."""
import numpy as np
import pytest
import torch

from oracle import cdr_oracle as O

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_RTOL, GRAD_ATOL = 1e-4, 1e-7


def dev():
    return torch.device('cuda', 0)


def ops():
    from recbole_cdr_b200 import ops as _ops
    return _ops


def lib():
    from recbole_cdr_b200 import _lib
    return _lib


def rand_table(n, d, seed, std=0.1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g) * std


def rand_ids(n, hi, seed, zipf=None):
    rng = np.random.RandomState(seed)
    if zipf:
        return torch.from_numpy(np.minimum(rng.zipf(zipf, n) - 1, hi - 1)).long()
    return torch.from_numpy(rng.randint(0, hi, n)).long()


def assert_grad_close(got, ref, what):
    # rows hit by thousands of duplicate ids are sums in a different order than index_add: allow 1e-4 of the
    # gradient's scale on top of the element-wise relative bound
    atol = max(GRAD_ATOL, 1e-4 * ref.abs().max().item())
    torch.testing.assert_close(got.cpu(), ref, rtol=GRAD_RTOL, atol=atol, msg=lambda m: f'{what}: {m}')


# ---------------------------------------------------------------------------------------------------- A1
@pytest.mark.parametrize('dim', [4, 32, 64, 96, 128, 160, 256])
@pytest.mark.parametrize('n_idx', [0, 1, 5, 1000])
def test_gather_rows_bit_exact(dim, n_idx):
    t = rand_table(777, dim, 1)
    idx = rand_ids(n_idx, 777, 2)
    out = ops().gather_rows_raw(t.to(dev()), idx.to(dev()))
    assert out.shape == (n_idx, dim)
    assert torch.equal(out.cpu(), O.gather_rows(t, idx))  # bit-exact


def test_gather_rows_into_concat_buffer_and_2d_index():
    t = rand_table(100, 64, 3).to(dev())
    idx = rand_ids(33, 100, 4).to(dev())
    buf = torch.zeros(33, 128, device=dev())
    ops().gather_rows_raw(t, idx, buf, 64)
    assert torch.equal(buf[:, 64:], t[idx]) and not buf[:, :64].any()
    out = ops().gather_rows(t, idx.view(-1, 1))  # the reference's [b, 1] overlap batch -> [b, 1, D]
    assert out.shape == (33, 1, 64) and torch.equal(out[:, 0], t[idx])


@pytest.mark.parametrize('dim', [64, 128, 36])
@pytest.mark.parametrize('zipf', [None, 1.05])
def test_scatter_add_matches_index_add(dim, zipf):
    n, b = 500, 4096
    idx = rand_ids(b, n, 5, zipf)
    rows = rand_table(b, dim, 6, 1.0)
    ref = torch.zeros(n, dim).index_add_(0, idx, rows * 0.5)
    dst = torch.zeros(n, dim, device=dev())
    ops().scatter_add_rows_raw(dst, idx.to(dev()), rows.to(dev()), 0.5)
    ref64 = torch.zeros(n, dim, dtype=torch.float64).index_add_(0, idx, rows.double() * 0.5)
    # fp32 sums of up to thousands of duplicates in atomic order vs index_add order: compare both to the fp64 sum
    tol = 1e-6 * torch.zeros(n, dim, dtype=torch.float64).index_add_(0, idx, rows.double().abs() * 0.5) + 1e-6
    assert ((dst.cpu().double() - ref64).abs() <= 8 * tol).all()
    assert ((ref.double() - ref64).abs() <= 8 * tol).all()


def test_scatter_add_all_duplicates():
    rows = torch.ones(8192, 64)
    dst = torch.zeros(4, 64, device=dev())
    ops().scatter_add_rows_raw(dst, torch.full((8192,), 2, dtype=torch.int64, device=dev()), rows.to(dev()))
    assert torch.equal(dst[2].cpu(), torch.full((64,), 8192.0)) and not dst[[0, 1, 3]].any()


def test_gather_autograd_backward_is_dense_index_add():
    t = rand_table(300, 64, 7).to(dev()).requires_grad_(True)
    idx = rand_ids(1000, 300, 8, 1.2)
    w = rand_table(1000, 64, 9, 1.0)
    (ops().gather_rows(t, idx.to(dev())) * w.to(dev())).sum().backward()
    ref = torch.zeros(300, 64, dtype=torch.float64).index_add_(0, idx, w.double())
    mag = torch.zeros(300, 64, dtype=torch.float64).index_add_(0, idx, w.double().abs())
    assert ((t.grad.cpu().double() - ref).abs() <= 1e-5 * mag + 1e-6).all()   # fp32 sum in atomic order vs fp64 sum


def test_out_of_range_id_raises_index_error():
    o = ops()
    t = rand_table(10, 64, 1).to(dev())
    old = o.CHECK_IDS
    o.CHECK_IDS = True
    try:
        with pytest.raises(IndexError):
            o.gather_rows_raw(t, torch.tensor([1, 10], device=dev()))
        with pytest.raises(IndexError):
            o.gather_rows_raw(t, torch.tensor([-1], device=dev()))
        o.gather_rows_raw(t, torch.tensor([9], device=dev()))  # flag was reset
    finally:
        o.CHECK_IDS = old


# ---------------------------------------------------------------------------------------------------- A2/A3
@pytest.mark.parametrize('batch,dim,nu,ni,zipf', [
    (1, 64, 50, 60, None), (3, 64, 50, 60, None), (257, 64, 1000, 1200, None), (8192, 64, 20000, 30000, None),
    (8192, 64, 20000, 30000, 1.05), (4096, 128, 5000, 5000, None), (1000, 32, 300, 300, 1.3), (513, 96, 300, 300, None),
    (2048, 256, 999, 999, None), (100, 36, 40, 40, None)])
def test_bpr_loss_and_grads(batch, dim, nu, ni, zipf):
    ut, it = rand_table(nu, dim, 11), rand_table(ni, dim, 12)
    u, ip, ineg = rand_ids(batch, nu, 13, zipf), rand_ids(batch, ni, 14, zipf), rand_ids(batch, ni, 15, zipf)
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    ref = O.emcdr_bpr_loss(a, b, u, ip, ineg, 0.01)
    gu_ref, gi_ref = O.grads_of(ref, [a, b])
    utc, itc = ut.to(dev()).requires_grad_(True), it.to(dev()).requires_grad_(True)
    loss = ops().bpr_loss(utc, itc, u.to(dev()), ip.to(dev()), ineg.to(dev()), 0.01)
    assert loss.shape == (1,)  # the reference's EmbLoss makes the loss shape [1]
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=LOSS_RTOL, atol=0)
    (loss * 1.7).backward()  # non-unit upstream gradient exercises the grad_loss path
    assert_grad_close(utc.grad, gu_ref * 1.7, 'user grad')
    assert_grad_close(itc.grad, gi_ref * 1.7, 'item grad')


def test_bpr_scores_are_dot_products():
    ut, it = rand_table(100, 64, 21), rand_table(120, 64, 22)
    u, i = rand_ids(77, 100, 23), rand_ids(77, 120, 24)
    s = ops().dot_score(ut.to(dev()), it.to(dev()), u.to(dev()), i.to(dev()))
    torch.testing.assert_close(s.cpu(), O.dot_score(ut, it, u, i), rtol=1e-5, atol=1e-6)


def test_bpr_loss_is_deterministic():
    ut, it = rand_table(5000, 64, 31).to(dev()), rand_table(5000, 64, 32).to(dev())
    u, ip, ineg = (rand_ids(8192, 5000, s).to(dev()) for s in (33, 34, 35))
    a = ops().bpr_loss(ut, it, u, ip, ineg, 0.01)
    for _ in range(5):
        assert torch.equal(a, ops().bpr_loss(ut, it, u, ip, ineg, 0.01))


@pytest.mark.parametrize('kind', ['mse', 'bce', 'none'])
@pytest.mark.parametrize('batch,dim,zipf', [(5, 64, None), (2048, 64, 1.05), (4097, 128, None), (300, 96, None)])
def test_point_loss_and_grads(kind, batch, dim, zipf):
    L = lib()
    nu, ni = 3000, 2000
    ut, it = rand_table(nu, dim, 41, 0.3), rand_table(ni, dim, 42, 0.3)
    u, i = rand_ids(batch, nu, 43, zipf), rand_ids(batch, ni, 44, zipf)
    y = (torch.rand(batch, generator=torch.Generator().manual_seed(45)) < 0.5).float()
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    if kind == 'mse':
        ref, k = O.emcdr_mf_loss(a, b, u, i, y, 0.01), L.LOSS_MSE
    elif kind == 'bce':
        ref = O.bce_loss(torch.sigmoid(O.dot_score(a, b, u, i)), y) + 0.01 * O.emb_loss(a[u], b[i])
        k = L.LOSS_BCE_SIGMOID
    else:
        ref, k = 0.01 * O.emb_loss(a[u], b[i]), L.LOSS_NONE
    gu_ref, gi_ref = O.grads_of(ref, [a, b])
    utc, itc = ut.to(dev()).requires_grad_(True), it.to(dev()).requires_grad_(True)
    loss = ops().point_loss(utc, itc, u.to(dev()), i.to(dev()), y.to(dev()), k, 0.01)
    torch.testing.assert_close(loss.cpu(), ref.detach().reshape(1), rtol=LOSS_RTOL, atol=0)
    loss.backward()
    assert_grad_close(utc.grad, gu_ref, 'user grad')
    assert_grad_close(itc.grad, gi_ref, 'item grad')


def test_inplace_table_grad_mode_matches_autograd_mode():
    o = ops()
    ut, it = rand_table(400, 64, 51), rand_table(400, 64, 52)
    u, ip, ineg = (rand_ids(1024, 400, s).to(dev()) for s in (53, 54, 55))
    res = {}
    for mode in ('autograd', 'inplace'):
        o.set_table_grad_mode(mode)
        try:
            a, b = ut.to(dev()).requires_grad_(True), it.to(dev()).requires_grad_(True)
            for _ in range(2):  # two backward passes accumulate, like two micro-batches
                o.bpr_loss(a, b, u, ip, ineg, 0.01).backward()
            res[mode] = (a.grad.clone(), b.grad.clone())
        finally:
            o.set_table_grad_mode('autograd')
    torch.testing.assert_close(res['inplace'][0], res['autograd'][0], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(res['inplace'][1], res['autograd'][1], rtol=1e-5, atol=1e-7)


def test_fused_sgd_scatter_updates_weights_in_place():
    """dst = the weight table itself, scale = -lr: the scatter IS the SGD step (xdr.h, xdr_bpr_bwd)."""
    L, o = lib(), ops()
    ut, it = rand_table(300, 64, 61), rand_table(300, 64, 62)
    u, ip, ineg = rand_ids(512, 300, 63), rand_ids(512, 300, 64), rand_ids(512, 300, 65)
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    gu, gi = O.grads_of(O.emcdr_bpr_loss(a, b, u, ip, ineg, 0.01), [a, b])
    lr = 0.05
    utc, itc = ut.to(dev()), it.to(dev())
    uc, pc, nc = u.to(dev()), ip.to(dev()), ineg.to(dev())
    scores = torch.empty(2, 512, device=dev())
    out8 = torch.empty(8, device=dev())
    s = L.cur_stream()
    L.call('xdr_bpr_fwd', utc.data_ptr(), itc.data_ptr(), 300, 300, 64, uc.data_ptr(), pc.data_ptr(), nc.data_ptr(), 512,
           1e-10, 0.01, scores[0].data_ptr(), scores[1].data_ptr(), out8.data_ptr(), L.workspace(dev()).data_ptr(), None, s)
    L.call('xdr_bpr_bwd', utc.data_ptr(), itc.data_ptr(), 300, 300, 64, uc.data_ptr(), pc.data_ptr(), nc.data_ptr(), 512,
           1e-10, 0.01, scores[0].data_ptr(), scores[1].data_ptr(), out8.data_ptr(), None, -lr, utc.data_ptr(),
           itc.data_ptr(), s)
    # rows hit several times read partially-updated values (Hogwild within a batch); at lr*grad ~ 1e-5 that is below atol
    torch.testing.assert_close(utc.cpu(), ut - lr * gu, rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(itc.cpu(), it - lr * gi, rtol=1e-4, atol=2e-6)


# ---------------------------------------------------------------------------------------------------- dense family
@pytest.mark.parametrize('M,N,K', [(1, 1, 8), (100, 64, 128), (8192, 128, 64), (1000, 8, 16), (777, 33, 20), (4096, 64, 256),
                                   (300, 24, 32), (257, 16, 64), (130, 128, 36)])   # (the last three: the 128 x 32 / 128 x 16 forward tiles, row tails)
@pytest.mark.parametrize('act', ['none', 'relu', 'tanh', 'sigmoid'])
def test_dense_fwd_bwd(M, N, K, act):
    L = lib()
    g = torch.Generator().manual_seed(71)
    X, W, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.2, torch.randn(N, generator=g) * 0.1
    G = torch.randn(M, N, generator=g)
    f = {'none': lambda v: v, 'relu': torch.relu, 'tanh': torch.tanh, 'sigmoid': torch.sigmoid}[act]
    xr, wr, br = (t.clone().requires_grad_(True) for t in (X, W, bias))
    yr = f(torch.nn.functional.linear(xr, wr, br))
    (yr * G).sum().backward()
    xc, wc, bc = (t.to(dev()).requires_grad_(True) for t in (X, W, bias))
    yc = ops().dense(xc, wc, bc, L.ACT_BY_NAME[act])
    # GEMM-shaped layers may run on the tcgen05 engine (tc5_dense.cu, when it is switched on): fp32 accumulation in tensor
    # memory rounds differently from an FMA chain -- a few 1e-5 absolute on sums of 64 .. 8192 terms (round 2, GPU call 13)
    big = M >= 128
    torch.testing.assert_close(yc.cpu(), yr.detach(), rtol=1e-4, atol=3e-5 if big else 1e-5)
    (yc * G.to(dev())).sum().backward()
    torch.testing.assert_close(xc.grad.cpu(), xr.grad, rtol=1e-4, atol=3e-5 if big else 1e-5)
    torch.testing.assert_close(wc.grad.cpu(), wr.grad, rtol=2e-4, atol=3e-4 if big else 1e-4)
    torch.testing.assert_close(bc.grad.cpu(), br.grad, rtol=2e-4, atol=3e-4 if big else 1e-4)


def test_dense_cross_stitch_unit():
    """relu(W x + b + m * (x2 H^T)) and its five gradients against autograd (conet.py:118-138)."""
    L = lib()
    g = torch.Generator().manual_seed(81)
    M, N, K = 3000, 64, 256
    X, X2 = torch.randn(M, K, generator=g), torch.randn(M, K, generator=g)
    W, H, bias = torch.randn(N, K, generator=g) * 0.1, torch.randn(N, K, generator=g) * 0.1, torch.randn(N, generator=g)
    ids = rand_ids(M, 100, 82)
    G = torch.randn(M, N, generator=g)
    ts = [t.clone().requires_grad_(True) for t in (X, W, bias, X2, H)]
    m = (ids < 40).float().unsqueeze(1)
    yr = torch.relu(torch.nn.functional.linear(ts[0], ts[1], ts[2]) + m * (ts[3] @ ts[4].t()))
    (yr * G).sum().backward()
    tc = [t.to(dev()).requires_grad_(True) for t in (X, W, bias, X2, H)]
    yc = ops().dense(tc[0], tc[1], tc[2], L.ACT_RELU, tc[3], tc[4], ids.to(dev()), 40)
    torch.testing.assert_close(yc.cpu(), yr.detach(), rtol=1e-4, atol=3e-5)   # (tensor-memory accumulation, see test_dense_fwd_bwd)
    (yc * G.to(dev())).sum().backward()
    for a, b, nm in zip(tc, ts, ('dX', 'dW', 'db', 'dX2', 'dH')):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=2e-4, atol=2e-4, msg=lambda s: f'{nm}: {s}')


def test_mse_rows_and_bce_logit():
    g = torch.Generator().manual_seed(91)
    T = rand_table(500, 64, 92, 0.5)
    idx = rand_ids(300, 500, 93, 1.2)
    Y = torch.randn(300, 64, generator=g)
    yr, tr = Y.clone().requires_grad_(True), T.clone().requires_grad_(True)
    ref = O.mse_loss(yr, tr[idx])
    ref.backward()
    yc, tc = Y.to(dev()).requires_grad_(True), T.to(dev()).requires_grad_(True)
    loss = ops().mse_rows(yc, tc, idx.to(dev()))
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=LOSS_RTOL, atol=0)
    loss.backward()
    torch.testing.assert_close(yc.grad.cpu(), yr.grad, rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(tc.grad.cpu(), tr.grad, rtol=1e-4, atol=1e-8)
    z = torch.randn(5000, generator=g) * 3
    z[0], z[1] = 200.0, -200.0  # saturated: exercises the log clamp at -100 and the 1e-12 eps of BCELoss backward
    y = (torch.rand(5000, generator=g) < 0.5).float()
    y[0], y[1] = 0.0, 1.0
    zr = z.clone().requires_grad_(True)
    ref = O.bce_loss(torch.sigmoid(zr), y)
    ref.backward()
    zc = z.to(dev()).requires_grad_(True)
    loss, prob = ops().bce_logit(zc, y.to(dev()))
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=LOSS_RTOL, atol=0)
    torch.testing.assert_close(prob.cpu(), torch.sigmoid(z), rtol=1e-6, atol=1e-7)
    loss.backward()
    torch.testing.assert_close(zc.grad.cpu(), zr.grad, rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize('shapes', [[(64, 256), (32, 64), (16, 32), (8, 16)], [(1, 1)], [(5, 3), (7, 1), (300, 333)],
                                    [(2, 2)] * 8])
def test_frob_sum(shapes):
    """CoNet's regulariser sum_l ||H_l||_F (conet.py:198-201) as one launch each way, against torch.norm per matrix: value,
    gradients under an upstream factor, and the all-zero matrix (torch.norm's gradient there is 0, not NaN)."""
    g = torch.Generator().manual_seed(len(shapes) * 17 + shapes[0][0])
    mats = [torch.randn(*sh, generator=g) * 0.3 for sh in shapes]
    if len(mats) > 1:
        mats[1].zero_()
    ref_in = [m.clone().requires_grad_(True) for m in mats]
    ref = sum(torch.norm(m) for m in ref_in)
    (ref * 0.37).backward()
    got_in = [m.to(dev()).requires_grad_(True) for m in mats]
    # the last matrix as a view that starts 4 bytes into its buffer: the scalar (not 16-byte aligned) read path
    pad = torch.zeros(mats[-1].numel() + 1, device=dev())
    pad[1:] = mats[-1].reshape(-1).to(dev())
    got_in[-1] = pad[1:].view(mats[-1].shape).detach().requires_grad_(True)
    assert got_in[-1].data_ptr() % 16 != 0 and got_in[-1].is_contiguous()
    got = ops().frob_sum(got_in)
    assert got.shape == ()
    (got * 0.37).backward()
    torch.testing.assert_close(got.detach().cpu(), ref.detach(), rtol=2e-6, atol=0)
    for a, b in zip(got_in, ref_in):
        assert bool(torch.isfinite(a.grad).all())
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-5, atol=1e-9)


def test_gather_max2_concat_fwd_bwd():
    su, tu, si, ti = (rand_table(200, 64, s) for s in (101, 102, 103, 104))
    tu[5] = su[5]  # exact ties: torch.maximum splits the gradient 0.5 / 0.5
    u, i = rand_ids(700, 200, 105, 1.3), rand_ids(700, 200, 106)
    u[:10] = 5
    G = rand_table(700, 128, 107, 1.0)
    r = [t.clone().requires_grad_(True) for t in (su, tu, si, ti)]
    ref = torch.cat((torch.maximum(r[0][u], r[1][u]), torch.maximum(r[2][i], r[3][i])), -1)
    (ref * G).sum().backward()
    c = [t.to(dev()).requires_grad_(True) for t in (su, tu, si, ti)]
    out = ops().GatherMax2Concat.apply(*c, u.to(dev()), i.to(dev()))
    assert torch.equal(out.cpu(), ref.detach())  # max of gathered rows is exact
    (out * G.to(dev())).sum().backward()
    for a, b in zip(c, r):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------------- fused MLP
@pytest.mark.parametrize('batch', [1, 31, 32, 33, 1000, 8192])
def test_fused_mlp_map_loss_matches_oracle(batch):
    """EMCDR map step at sizes around the 32-row tile (incl. ragged tiles and a single row)."""
    g = torch.Generator().manual_seed(111)
    src, tgt = rand_table(3000, 64, 112, 0.3), rand_table(3000, 64, 113, 0.3)
    ws = [torch.randn(128, 64, generator=g) * 0.2, torch.randn(64, 128, generator=g) * 0.2]
    bs = [torch.randn(128, generator=g) * 0.1, torch.randn(64, generator=g) * 0.1]
    idx = rand_ids(batch, 3000, 114, 1.3)
    leaves = [t.clone().requires_grad_(True) for t in [src, tgt] + ws + bs]
    ref = O.emcdr_map_loss(leaves[0], leaves[1], idx.view(-1, 1), leaves[2:4], leaves[4:6])
    ref.backward()
    c = [t.to(dev()).requires_grad_(True) for t in [src, tgt] + ws + bs]
    assert ops().fused_mlp_supported([64, 128, 64])
    loss = ops().fused_mlp_loss(0, 0, lib().ACT_TANH, idx.to(dev()), None, None, (c[0], None, None, None, c[1]), c[2:4], c[4:6])
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=LOSS_RTOL, atol=0)
    (loss * 1.3).backward()
    for got, want, nm in zip(c, leaves, ('src', 'tgt', 'W1', 'W2', 'b1', 'b2')):
        atol = max(1e-7, 1e-4 * want.grad.abs().max().item())
        torch.testing.assert_close(got.grad.cpu(), want.grad * 1.3, rtol=2e-4, atol=atol, msg=lambda s: f'{nm}: {s}')


def test_fused_mlp_unsupported_stacks_fall_back():
    assert not ops().fused_mlp_supported([512, 64, 1])       # wider than 256
    assert not ops().fused_mlp_supported([256, 256, 256])     # more than 32 weight elements per thread
    assert ops().fused_mlp_supported([128, 32, 16, 1])
