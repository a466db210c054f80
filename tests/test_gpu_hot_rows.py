"""GPU: hot rows (xdr_steps_set_hot_rows) -- the gradients of popular rows are pre-aggregated per CTA in shared memory and
flushed once per launch.  Same losses bit for bit, same gradient tables up to summation order, on Zipf ids at full batch size."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('pairwise,dim', [(True, 64), (False, 64), (True, 32)])
def test_hot_rows_give_the_same_gradients(pairwise, dim):
    from recbole_cdr_b200 import _lib, ops
    dev = torch.device('cuda', 0)
    K, B, nu, ni = 12, 8192, 50_000, 70_000
    g = torch.Generator().manual_seed(5)
    rng = np.random.RandomState(5)
    ut, it = (torch.randn(nu, dim, generator=g) * 0.1).to(dev), (torch.randn(ni, dim, generator=g) * 0.1).to(dev)
    u = torch.from_numpy(rng.randint(1, nu, (K, B))).long().to(dev)
    ia = torch.from_numpy(np.minimum(rng.zipf(1.05, (K, B)), ni - 1)).long().to(dev)
    ib = torch.from_numpy(np.minimum(rng.zipf(1.05, (K, B)), ni - 1)).long().to(dev)
    y = (torch.rand(K, B, generator=g) < 0.5).float().to(dev)
    args = (ut, it, u, ia, ib) if pairwise else (ut, it, u, ia, None, y)
    kw = dict(reg_weight=0.01) if pairwise else dict(reg_weight=0.01, loss_kind=_lib.LOSS_BCE_SIGMOID)
    o0, gu0, gi0 = ops.train_steps(*args, **kw)
    hot_i = ops.hot_rows_from_ids(torch.stack([ia, ib]), ni, 48)
    hot_u = ops.hot_rows_from_ids(u, nu, 16)
    assert hot_i.numel() >= 10 and hot_u.numel() == 0     # Zipf items have popular rows, uniform users have none
    ops.set_steps_hot_rows(hot_u, torch.cat([hot_i, torch.tensor([ni + 5, int(hot_i[0])], device=dev)]))   # + a bad id, a duplicate
    try:
        o1, gu1, gi1 = ops.train_steps(*args, **kw)
    finally:
        ops.set_steps_hot_rows(None, None)
    torch.cuda.synchronize()
    assert torch.equal(o0[:, 0], o1[:, 0])
    torch.testing.assert_close(gu1, gu0, rtol=1e-5, atol=1e-5 * float(gu0.abs().max()))
    torch.testing.assert_close(gi1, gi0, rtol=1e-5, atol=1e-4 * float(gi0.abs().max()))
