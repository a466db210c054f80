"""GPU: lazily zeroed gradient tables (xdr_train_steps_lazy, include/xdr.h).  The first touch of a destination row in a launch
zero-fills it (no DRAM read of gradient lines), the scatter-adds wait for the row's "filled" bit.  Checked against the plain
scatter-add launch of the same kernel (hardware-validated in round 1) and against the oracle, on garbage-initialised
destinations, with heavy duplication (small tables, Zipf ids) and at BASELINE configs[1]'s full size."""
import numpy as np
import pytest
import torch

from oracle import cdr_oracle as O

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda', 0)


def make(nu, ni, dim, K, B, seed, zipf=None):
    g = torch.Generator().manual_seed(seed)
    ut, it = (torch.randn(nu, dim, generator=g) * 0.1).to(dev()), (torch.randn(ni, dim, generator=g) * 0.1).to(dev())
    rng = np.random.RandomState(seed)

    def ids(hi):
        if zipf:
            return torch.from_numpy(np.minimum(rng.zipf(zipf, (K, B)), hi - 1)).long().to(dev())
        return torch.from_numpy(rng.randint(1, hi, (K, B))).long().to(dev())
    return ut, it, ids(nu), ids(ni), ids(ni)


def named(n, *ids):
    m = torch.zeros(n, dtype=torch.bool, device=dev())
    for i in ids:
        m[i.reshape(-1)] = True
    return m


@pytest.mark.parametrize('nu,ni,dim,K,B,zipf', [(300, 400, 64, 12, 8192, None), (5000, 7000, 64, 40, 8192, None),
                                                (5000, 7000, 64, 25, 8192, 1.2), (2000, 3000, 32, 9, 4096, None),
                                                (1_500_001, 2_000_001, 64, 20, 8192, None)])
def test_lazy_tables_equal_the_plain_scatter_add(nu, ni, dim, K, B, zipf):
    from recbole_cdr_b200 import ops
    ut, it, u, ip, ineg = make(nu, ni, dim, K, B, 3, zipf)
    o_ref, gu_ref, gi_ref = ops.train_steps(ut, it, u, ip, ineg, reg_weight=0.01)
    tm = ops.TouchMap(nu, ni, dev())
    tm.words.fill_(-1)   # stale marks: fresh=True clears them
    gu, gi = torch.full_like(ut, 7.0), torch.full_like(it, -3.0)      # garbage: untouched rows must stay garbage
    o, _, _ = ops.train_steps(ut, it, u, ip, ineg, reg_weight=0.01, user_dst=gu, item_dst=gi, touch=tm, fresh=True)
    torch.cuda.synchronize()
    assert torch.equal(o[:, 0], o_ref[:, 0])                         # the loss path is untouched
    su, si = tm.state()
    nu_m, ni_m = named(nu, u), named(ni, ip, ineg)
    assert torch.equal(su != 0, nu_m) and torch.equal(si != 0, ni_m)
    assert bool(((su == 0) | (su == 3)).all()) and bool(((si == 0) | (si == 3)).all())
    # rows named thousands of times (the Zipf case: 33 000 duplicates of one row) are summed in another order than by the
    # plain launch: 4e-8 absolute on a 0.018 row was measured between two correct runs (scripts/diag_accum_noise.py)
    tol = 1e-5 if zipf else 1e-6
    atol_u, atol_i = tol * float(gu_ref.abs().max()), tol * float(gi_ref.abs().max())
    torch.testing.assert_close(gu[nu_m], gu_ref[nu_m], rtol=1e-5, atol=max(atol_u, 1e-9))
    torch.testing.assert_close(gi[ni_m], gi_ref[ni_m], rtol=1e-5, atol=max(atol_i, 1e-9))
    assert bool((gu[~nu_m] == 7.0).all()) and bool((gi[~ni_m] == -3.0).all())
    # a second launch without clearing keeps accumulating: twice the gradient on the marked rows.  Tolerance: adding a row's
    # contributions a second time ON TOP of their sum rounds differently from doubling the sum -- measured on a B200
    # (scripts/diag_accum_noise.py, Zipf ids, 33 000 duplicates of one row): 1.8e-6 absolute = 1e-4 of the table's maximum,
    # identically for plain and lazily zeroed destinations.
    ops.train_steps(ut, it, u, ip, ineg, reg_weight=0.01, user_dst=gu, item_dst=gi, touch=tm, fresh=False)
    torch.cuda.synchronize()
    torch.testing.assert_close(gu[nu_m], 2 * gu_ref[nu_m], rtol=1e-5, atol=max(2e-4 * float(gu_ref.abs().max()), 1e-9))
    torch.testing.assert_close(gi[ni_m], 2 * gi_ref[ni_m], rtol=1e-5, atol=max(2e-4 * float(gi_ref.abs().max()), 1e-9))


def test_lazy_tables_match_the_oracle_and_pointwise_kinds():
    from recbole_cdr_b200 import _lib, ops
    nu, ni, dim, K, B = 900, 1100, 64, 5, 2048
    ut, it, u, i, _ = make(nu, ni, dim, K, B, 9)
    y = (torch.rand(K, B, generator=torch.Generator().manual_seed(1)) < 0.5).float().to(dev())
    tm = ops.TouchMap(nu, ni, dev())
    o, gu, gi = ops.train_steps(ut, it, u, i, None, y, loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=0.01,
                                user_dst=torch.full_like(ut, 2.0), item_dst=torch.full_like(it, 2.0), touch=tm, fresh=True)
    torch.cuda.synchronize()
    a, b = ut.cpu().requires_grad_(True), it.cpu().requires_grad_(True)
    total = 0
    for s in range(K):
        ref = O.bce_loss(torch.sigmoid(O.dot_score(a, b, u[s].cpu(), i[s].cpu())), y[s].cpu()) + \
            0.01 * O.emb_loss(a[u[s].cpu()], b[i[s].cpu()])
        torch.testing.assert_close(o[s, 0].cpu(), ref.detach().reshape(-1)[0], rtol=1e-4, atol=0)
        total = total + ref.sum()
    du, di = torch.autograd.grad(total, [a, b])
    tu, ti = tm.touched()
    torch.testing.assert_close(gu[tu].cpu(), du[tu.cpu()], rtol=1e-4, atol=1e-4 * float(du.abs().max()))
    torch.testing.assert_close(gi[ti].cpu(), di[ti.cpu()], rtol=1e-4, atol=1e-4 * float(di.abs().max()))
    assert bool((gu[~tu] == 2.0).all())
    with pytest.raises(_lib.XdrError, match='cannot be the weight table'):
        ops.train_steps(ut, it, u, i, None, y, loss_kind=_lib.LOSS_BCE_SIGMOID, user_dst=ut, item_dst=it, touch=tm)
