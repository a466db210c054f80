"""Size-independent properties of the training hot path (gather -> BPR score + loss -> gradient scatter-add), written once
and run (a) at BASELINE.json's full size on the GPU (tests/test_gpu_z_fullsize.py: 1M x 1M rows, dim 64, B = 8192) where a
CPU oracle pass would take minutes, and (b) at toy size through the CTA emulator (tests/test_emu_fullsize_props.py), which
is how the checker itself was tested while no GPU was reachable.

All of them follow from the reference's arithmetic (emcdr.py:110-132, BPRLoss + reg_weight * EmbLoss):
  P1 support      a gradient row is non-zero only if the batch names it; tables are otherwise untouched
  P2 item balance with reg_weight = 0 the positive item gets c*Eu and the negative item -c*Eu: every column of the item
                  gradient table sums to zero over the rows (up to fp32 rounding of the scatter)
  P3 user sum     the column sums of the user gradient table equal sum_b c_b (Ei+ - Ei-), computed here from the ids
  P4 linearity    scale = 2 doubles every gradient row (bit-exact up to the order of duplicate-row additions)
  P5 additivity   K steps in one launch accumulate what K single-step launches accumulate, and the per-step losses are the
                  same bits (the loss of a batch does not depend on its neighbours in the launch)
  P6 checksum     sum of per-step losses equals the oracle's (plain torch on the same device) within 1e-4 relative
"""
import torch


def _ids(rng, hi, K, B, device):
    return torch.from_numpy(rng.randint(1, hi, (K, B))).long().to(device)


def bpr_coefficients(ut, it, u, ip, ineg, gamma=1e-10):
    """d loss / d (pos - neg) per interaction for BPRLoss = -mean(log(gamma + sigmoid(pos - neg))): plain torch, any device."""
    eu, ep, en = ut[u], it[ip], it[ineg]
    x = ((eu * ep).sum(1) - (eu * en).sum(1)).double()
    s = torch.sigmoid(x)
    c = -(s * (1 - s)) / (gamma + s) / x.numel()
    return c, eu.double(), ep.double(), en.double()


def check_train_step_properties(ops, device, nu, ni, dim, K, B, seed=0, oracle=None):
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    ut = (torch.randn(nu, dim, generator=g) * 0.1).to(device)
    it = (torch.randn(ni, dim, generator=g) * 0.1).to(device)
    rng = np.random.RandomState(seed)
    u, ip, ineg = _ids(rng, nu, K, B, device), _ids(rng, ni, K, B, device), _ids(rng, ni, K, B, device)
    ut0, it0 = ut.clone(), it.clone()

    out8, gu, gi = ops.train_steps(ut, it, u, ip, ineg, reg_weight=0.0)
    assert torch.equal(ut, ut0) and torch.equal(it, it0), 'the weight tables must not change in gradient-table mode'

    # P1 support
    named_u = torch.zeros(nu, dtype=torch.bool, device=device)
    named_u[u.reshape(-1)] = True
    named_i = torch.zeros(ni, dtype=torch.bool, device=device)
    named_i[ip.reshape(-1)] = True
    named_i[ineg.reshape(-1)] = True
    assert not bool(gu[~named_u].any()), 'gradient on a user row no batch names'
    assert not bool(gi[~named_i].any()), 'gradient on an item row no batch names'
    assert int((gu.abs().sum(1) > 0).sum()) >= 0.99 * int(named_u.sum())

    # P2 item balance, P3 user column sums (fp64 accumulation of the fp32 tables)
    col_i = gi.double().sum(0)
    ref_u = torch.zeros(dim, dtype=torch.float64, device=device)
    mass = 0.0
    for k in range(K):
        c, eu, ep, en = bpr_coefficients(ut, it, u[k], ip[k], ineg[k])
        ref_u += (c.unsqueeze(1) * (ep - en)).sum(0)
        mass += float((c.abs().unsqueeze(1) * eu.abs()).sum())
    # every scatter-add rounds to fp32 once: allow 2^-22 of the absolute mass that went into the columns
    assert float(col_i.abs().max()) <= 4 * 2.0 ** -22 * mass, (float(col_i.abs().max()), mass)
    torch.testing.assert_close(gu.double().sum(0), ref_u, rtol=1e-4, atol=1e-4 * float(ref_u.abs().max()))

    # P4 linearity in `scale` (x2 is exact in fp32; only the order of duplicate-row additions may differ)
    _, gu2, gi2 = ops.train_steps(ut, it, u, ip, ineg, reg_weight=0.0, scale=2.0)
    torch.testing.assert_close(gu2, 2 * gu, rtol=1e-5, atol=1e-6 * float(gu.abs().max()))
    torch.testing.assert_close(gi2, 2 * gi, rtol=1e-5, atol=1e-6 * float(gi.abs().max()))

    # P5 additivity over launches + per-step losses independent of the launch they ran in
    gu_acc, gi_acc = torch.zeros_like(ut), torch.zeros_like(it)
    k_split = max(1, K // 2)
    o_a, _, _ = ops.train_steps(ut, it, u[:k_split], ip[:k_split], ineg[:k_split], reg_weight=0.0, user_dst=gu_acc,
                                item_dst=gi_acc)
    loss_a = o_a[:, 0].clone()
    if k_split < K:
        o_b, _, _ = ops.train_steps(ut, it, u[k_split:], ip[k_split:], ineg[k_split:], reg_weight=0.0, user_dst=gu_acc,
                                    item_dst=gi_acc)
        loss_ab = torch.cat([loss_a, o_b[:, 0]])
    else:
        loss_ab = loss_a
    assert torch.equal(loss_ab, out8[:, 0]), 'per-step losses must not depend on how the steps are grouped into launches'
    torch.testing.assert_close(gu_acc, gu, rtol=1e-5, atol=1e-6 * float(gu.abs().max()))
    torch.testing.assert_close(gi_acc, gi, rtol=1e-5, atol=1e-6 * float(gi.abs().max()))

    # P6 checksum of the per-step losses against the oracle (plain torch ops on the same device), with reg_weight on
    if oracle is not None:
        out8r, gur, gir = ops.train_steps(ut, it, u, ip, ineg, reg_weight=0.01)
        a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
        total = 0
        for k in range(K):
            ref = oracle.emcdr_bpr_loss(a, b, u[k], ip[k], ineg[k], 0.01)
            torch.testing.assert_close(out8r[k, 0], ref.detach().reshape(-1)[0], rtol=1e-4, atol=0)
            total = total + ref.sum()
        du, di = torch.autograd.grad(total, [a, b])
        torch.testing.assert_close(gur, du, rtol=1e-4, atol=1e-4 * float(du.abs().max()))
        torch.testing.assert_close(gir, di, rtol=1e-4, atol=1e-4 * float(di.abs().max()))
    return out8


def check_spmm_properties(NormAdj, device, nu, ni, n_edges, dim, seed=0, zipf=None):
    """BiTGCF's normalised adjacency L = D^-1/2 A D^-1/2 (bitgcf.py:92-116) and its SpMM, at any size:
      G1 row sums   (L 1)_v = sum_{w in N(v)} (d_v + 1e-7)^-1/2 (d_w + 1e-7)^-1/2, computed here in fp64 from the edge list
      G2 symmetry   <Y, L X> = <L Y, X>  (the backward pass of the propagation relies on it)
      G3 linearity  L (2 X + Y) = 2 L X + L Y
      G4 support    rows of isolated nodes are exactly zero
    """
    import numpy as np
    rng = np.random.RandomState(seed)
    r = rng.randint(0, nu, n_edges)
    c = np.minimum(rng.zipf(zipf, n_edges) - 1, ni - 1) if zipf else rng.randint(0, ni, n_edges)
    adj = NormAdj(r, c, nu, ni, device)
    n = nu + ni
    # the reference de-duplicates edges (dict keys) and symmetrises: node ids are users [0, nu) then items [nu, nu + ni)
    e = np.unique(r.astype(np.int64) * ni + c)
    er, ec = e // ni, e % ni + nu
    deg = np.bincount(er, minlength=n) + np.bincount(ec, minlength=n)
    dinv = np.power(deg + 1e-7, -0.5)
    w = dinv[er] * dinv[ec]
    rowsum = np.bincount(er, weights=w, minlength=n) + np.bincount(ec, weights=w, minlength=n)

    g = torch.Generator().manual_seed(seed)
    ones = torch.ones(n, dim, device=device)
    got = adj.spmm(ones)
    ref = torch.from_numpy(rowsum).to(device)
    torch.testing.assert_close(got[:, 0].double(), ref, rtol=2e-4, atol=1e-6)
    # (split rows accumulate their partial sums with atomics, whose order may differ from column to column: not bit-equal)
    torch.testing.assert_close(got[:, 0], got[:, dim - 1], rtol=1e-5, atol=1e-7)
    isolated = torch.from_numpy(deg == 0).to(device)
    assert not bool(got[isolated].any())                                                  # G4

    X = torch.randn(n, dim, generator=g).to(device)
    Y = torch.randn(n, dim, generator=g).to(device)
    LX, LY = adj.spmm(X), adj.spmm(Y)
    a, b = (Y.double() * LX.double()).sum(), (LY.double() * X.double()).sum()
    scale = float((Y.double().abs() * LX.double().abs()).sum())
    assert abs(float(a - b)) <= 1e-6 * scale, (float(a), float(b), scale)                 # G2
    L2 = adj.spmm(2 * X + Y)
    torch.testing.assert_close(L2, 2 * LX + LY, rtol=1e-4, atol=1e-5)                     # G3
    return adj
