"""The persistent multi-step trainer kernel (steps_persistent.cu -- the headline kernel, parity-green on a B200) through the
CPU CTA emulator with ALL of its CTAs resident at once: TMA id-tile ring, mbarrier hand-offs between producer / loader /
scatter warps and the grid-wide tagged-word norm exchange run on the emulator's mbarrier / bulk-copy model.  Pins that
model against a hardware-validated kernel and keeps the headline path's logic checkable on CPU."""
import numpy as np
import pytest
import torch

import emu_util
from oracle import cdr_oracle as O


def setup(nu, ni, dim, K, B, seed, std=0.1):
    g = torch.Generator().manual_seed(seed)
    ut, it = torch.randn(nu, dim, generator=g) * std, torch.randn(ni, dim, generator=g) * std
    rng = np.random.RandomState(seed)
    ids = lambda hi: torch.from_numpy(rng.randint(0, hi, (K, B))).long()
    return ut, it, ids(nu), ids(ni), ids(ni), (torch.rand(K, B, generator=g) < 0.5).float()


@pytest.mark.parametrize('K,B,dim,sms,seed', [(3, 96, 64, 3, 0), (5, 64, 64, 2, 0), (2, 40, 32, 4, 0), (4, 128, 128, 2, 0),
                                              (3, 96, 64, 3, 17)])
def test_train_steps_bpr_matches_oracle_per_step(K, B, dim, sms, seed):
    nu, ni = 300, 400
    ut, it, u, ip, ineg, _ = setup(nu, ni, dim, K, B, 5)
    with emu_util.patched_ops(sms=sms, seed=seed) as ops:
        out8, gu, gi = ops.train_steps(ut.clone(), it.clone(), u, ip, ineg, reg_weight=0.01)
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    gu_ref, gi_ref = torch.zeros_like(ut), torch.zeros_like(it)
    for k in range(K):
        ref = O.emcdr_bpr_loss(a, b, u[k], ip[k], ineg[k], 0.01)
        torch.testing.assert_close(out8[k, 0], ref.detach()[0], rtol=1e-4, atol=0)
        du, di = O.grads_of(ref, [a, b])
        gu_ref += du
        gi_ref += di
    torch.testing.assert_close(gu, gu_ref, rtol=1e-4, atol=1e-4 * gu_ref.abs().max().item())
    torch.testing.assert_close(gi, gi_ref, rtol=1e-4, atol=1e-4 * gi_ref.abs().max().item())


@pytest.mark.parametrize('kind', ['mse', 'bce'])
def test_train_steps_pointwise(kind):
    from recbole_cdr_b200 import _lib
    K, B, dim, nu, ni = 3, 64, 64, 200, 250
    ut, it, u, i, _, y = setup(nu, ni, dim, K, B, 7, 0.3)
    k = _lib.LOSS_MSE if kind == 'mse' else _lib.LOSS_BCE_SIGMOID
    with emu_util.patched_ops(sms=2) as ops:
        out8, gu, gi = ops.train_steps(ut.clone(), it.clone(), u, i, None, y, loss_kind=k, reg_weight=0.01)
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    gu_ref, gi_ref = torch.zeros_like(ut), torch.zeros_like(it)
    for s in range(K):
        ref = (O.emcdr_mf_loss(a, b, u[s], i[s], y[s], 0.01) if kind == 'mse' else
               O.bce_loss(torch.sigmoid(O.dot_score(a, b, u[s], i[s])), y[s]) + 0.01 * O.emb_loss(a[u[s]], b[i[s]]))
        torch.testing.assert_close(out8[s, 0], ref.detach().reshape(-1)[0], rtol=1e-4, atol=0)
        du, di = O.grads_of(ref, [a, b])
        gu_ref += du
        gi_ref += di
    torch.testing.assert_close(gu, gu_ref, rtol=1e-4, atol=1e-4 * gu_ref.abs().max().item())
    torch.testing.assert_close(gi, gi_ref, rtol=1e-4, atol=1e-4 * gi_ref.abs().max().item())


@pytest.mark.parametrize('world', [2, 4])
def test_row_sharded_steps_equal_the_unsharded_oracle(world):
    """E1 on one host: the tables are split block-cyclically into ``world`` shards (row r -> shard r mod G, local row
    r div G), every "rank" runs its own batches through xdr_train_steps_sharded against ALL shards (host pointers stand in
    for the CUDA-IPC peer mappings), gradients land in the owners' shards.  Per-batch losses and the re-assembled gradient
    tables must equal the oracle over the union of the batches."""
    from recbole_cdr_b200.shard import RowShardedTable, train_steps_sharded
    K, B, dim, nu, ni = 2, 64, 64, 301, 403
    ut, it, u, ip, ineg, _ = setup(nu, ni, dim, K * world, B, 11)
    with emu_util.patched_ops(sms=2):
        def shards(full):
            tabs = [RowShardedTable.from_full(full, r, world, 'cpu') for r in range(world)]
            for t in tabs:
                t._ptrs = [s.local.data_ptr() for s in tabs]
            return tabs
        tu, ti = shards(ut), shards(it)
        gu, gi = shards(torch.zeros_like(ut)), shards(torch.zeros_like(it))
        losses = []
        for r in range(world):
            sl = slice(r * K, (r + 1) * K)
            out8 = train_steps_sharded(tu[r], ti[r], gu[r], gi[r], u[sl].contiguous(), ip[sl].contiguous(),
                                       ineg[sl].contiguous(), reg_weight=0.01)
            losses.append(out8[:, 0].clone())

    def assemble(tabs, n_rows):
        full = torch.empty((tabs[0].local.shape[0] * world, dim))
        for r, t in enumerate(tabs):
            full[r::world] = t.local
        return full[:n_rows]

    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    gu_ref, gi_ref = torch.zeros_like(ut), torch.zeros_like(it)
    for k in range(K * world):
        ref = O.emcdr_bpr_loss(a, b, u[k], ip[k], ineg[k], 0.01)
        torch.testing.assert_close(torch.cat(losses)[k], ref.detach()[0], rtol=1e-4, atol=0)
        du, di = O.grads_of(ref, [a, b])
        gu_ref += du
        gi_ref += di
    torch.testing.assert_close(assemble(gu, nu), gu_ref, rtol=1e-4, atol=1e-4 * gu_ref.abs().max().item())
    torch.testing.assert_close(assemble(gi, ni), gi_ref, rtol=1e-4, atol=1e-4 * gi_ref.abs().max().item())
    assert torch.equal(assemble(tu, nu), ut)      # the weight shards are only read
