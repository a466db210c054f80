"""GPU parity of the persistent multi-step kernel (xdr_train_steps) against the per-step kernels and the oracle."""
import numpy as np
import pytest
import torch

from oracle import cdr_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda', 0)


def setup(nu, ni, dim, K, B, seed, zipf=None, std=0.1):
    g = torch.Generator().manual_seed(seed)
    ut, it = torch.randn(nu, dim, generator=g) * std, torch.randn(ni, dim, generator=g) * std
    rng = np.random.RandomState(seed)

    def ids(hi):
        if zipf:
            return torch.from_numpy(np.minimum(rng.zipf(zipf, (K, B)) - 1, hi - 1)).long()
        return torch.from_numpy(rng.randint(0, hi, (K, B))).long()

    return ut, it, ids(nu), ids(ni), ids(ni), (torch.rand(K, B, generator=g) < 0.5).float()


@pytest.mark.parametrize('K,B,dim,zipf', [(1, 8192, 64, None), (7, 8192, 64, None), (5, 4096, 64, 1.1), (9, 256, 64, None),
                                          (3, 4, 64, None), (6, 1000, 32, None), (4, 2048, 128, None), (3, 512, 96, None),
                                          (40, 8192, 64, None), (3, 16384, 64, None), (20, 148 * 4, 64, None), (4, 23680, 64, None), (12, 8192, 128, None),
                                          (5, 2052, 256, None), (33, 36, 16, None)])
def test_train_steps_bpr_matches_oracle_per_step(K, B, dim, zipf):
    from recbole_cdr_b200 import ops
    nu, ni = 3000, 4000
    ut, it, u, ip, ineg, _ = setup(nu, ni, dim, K, B, 5, zipf)
    assert ops.train_steps_supported(B, dim, True)
    out8, gu, gi = ops.train_steps(ut.to(dev()), it.to(dev()), u.to(dev()), ip.to(dev()), ineg.to(dev()), reg_weight=0.01)
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    gu_ref, gi_ref = torch.zeros_like(ut), torch.zeros_like(it)
    for k in range(K):
        ref = O.emcdr_bpr_loss(a, b, u[k], ip[k], ineg[k], 0.01)
        torch.testing.assert_close(out8[k, 0].cpu(), ref.detach()[0], rtol=1e-4, atol=0)
        du, di = O.grads_of(ref, [a, b])
        gu_ref += du
        gi_ref += di
    scale_u, scale_i = gu_ref.abs().max().item(), gi_ref.abs().max().item()
    torch.testing.assert_close(gu.cpu(), gu_ref, rtol=1e-4, atol=1e-4 * scale_u)
    torch.testing.assert_close(gi.cpu(), gi_ref, rtol=1e-4, atol=1e-4 * scale_i)


@pytest.mark.parametrize('kind', ['mse', 'bce'])
def test_train_steps_pointwise(kind):
    from recbole_cdr_b200 import _lib, ops
    K, B, dim, nu, ni = 5, 2048, 64, 2000, 2500
    ut, it, u, i, _, y = setup(nu, ni, dim, K, B, 7, None, 0.3)
    k = _lib.LOSS_MSE if kind == 'mse' else _lib.LOSS_BCE_SIGMOID
    out8, gu, gi = ops.train_steps(ut.to(dev()), it.to(dev()), u.to(dev()), i.to(dev()), None, y.to(dev()), loss_kind=k,
                                   reg_weight=0.01)
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    gu_ref, gi_ref = torch.zeros_like(ut), torch.zeros_like(it)
    for s in range(K):
        if kind == 'mse':
            ref = O.emcdr_mf_loss(a, b, u[s], i[s], y[s], 0.01)
        else:
            ref = O.bce_loss(torch.sigmoid(O.dot_score(a, b, u[s], i[s])), y[s]) + 0.01 * O.emb_loss(a[u[s]], b[i[s]])
        torch.testing.assert_close(out8[s, 0].cpu(), ref.detach().reshape(()), rtol=1e-4, atol=0)
        du, di = O.grads_of(ref, [a, b])
        gu_ref += du
        gi_ref += di
    torch.testing.assert_close(gu.cpu(), gu_ref, rtol=1e-4, atol=1e-4 * gu_ref.abs().max().item())
    torch.testing.assert_close(gi.cpu(), gi_ref, rtol=1e-4, atol=1e-4 * gi_ref.abs().max().item())


@pytest.mark.parametrize('force_regs', [0, 1])
def test_train_steps_equals_per_step_kernels_and_is_deterministic_in_loss(force_regs):
    import ctypes
    from recbole_cdr_b200 import _lib, ops
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.xdr_debug_force_register_kernel(force_regs)  # exercise both persistent kernels on the headline shape
    try:
        _check_equals_per_step(ops)
    finally:
        lib.xdr_debug_force_register_kernel(0)


def _check_equals_per_step(ops):
    K, B, dim, nu, ni = 12, 8192, 64, 50000, 60000
    ut, it, u, ip, ineg, _ = setup(nu, ni, dim, K, B, 11)
    utc, itc, uc, pc, nc = (t.to(dev()) for t in (ut, it, u, ip, ineg))
    out_a, gu_a, gi_a = ops.train_steps(utc, itc, uc, pc, nc, reg_weight=0.01)
    out_b, _, _ = ops.train_steps(utc, itc, uc, pc, nc, reg_weight=0.01)
    assert torch.equal(out_a, out_b)  # fixed-order reductions: bitwise reproducible losses
    a, b = utc.clone().requires_grad_(True), itc.clone().requires_grad_(True)
    for k in range(K):
        loss = ops.bpr_loss(a, b, uc[k], pc[k], nc[k], 0.01)
        torch.testing.assert_close(out_a[k, 0:1], loss.detach(), rtol=2e-6, atol=0)
        loss.backward()
    torch.testing.assert_close(gu_a, a.grad, rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(gi_a, b.grad, rtol=1e-4, atol=1e-9)


def test_train_steps_strided_id_buffer_and_unsupported_shapes():
    from recbole_cdr_b200 import _lib, ops
    K, B = 4, 1024
    ut, it, u, ip, ineg, _ = setup(500, 500, 64, K, B, 13)
    packed = torch.stack([u, ip, ineg], dim=1).to(dev())  # [K, 3, B]: the layout the trainer's H2D copy produces
    out8, gu, gi = ops.train_steps(ut.to(dev()), it.to(dev()), packed[:, 0], packed[:, 1], packed[:, 2], reg_weight=0.01)
    ref8, gu2, gi2 = ops.train_steps(ut.to(dev()), it.to(dev()), u.to(dev()), ip.to(dev()), ineg.to(dev()), reg_weight=0.01)
    assert torch.equal(out8, ref8)
    torch.testing.assert_close(gu, gu2, rtol=1e-5, atol=1e-9)
    assert ops.train_steps_supported(16384, 64, True)
    assert not ops.train_steps_supported(65536, 64, True)      # more than 2 tasks per worker per step: per-step kernels
    assert not ops.train_steps_supported(8190, 64, True)       # batch % 4 != 0 (TMA id tiles need 16-byte alignment)
    odd = torch.zeros(1, 8190, dtype=torch.int64, device=dev())
    with pytest.raises(_lib.XdrError, match='per-step'):
        ops.train_steps(ut.to(dev()), it.to(dev()), odd, odd, odd)


def test_fused_sgd_steps_train():
    """dst = the tables, scale = -lr: K persistent steps of asynchronous SGD lower the BPR loss on a fixed batch set."""
    from recbole_cdr_b200 import ops
    K, B = 64, 2048
    ut, it, u, ip, ineg, _ = setup(800, 900, 64, 4, B, 17)
    utc, itc = ut.to(dev()), it.to(dev())
    rep = lambda t: t.to(dev()).repeat(K // 4, 1)
    out8, _, _ = ops.train_steps(utc, itc, rep(u), rep(ip), rep(ineg), reg_weight=0.0, user_dst=utc, item_dst=itc,
                                 scale=-20.0)
    l = out8[:, 0].cpu()
    assert torch.isfinite(l).all() and l[-4:].mean() < l[:4].mean() - 0.05
