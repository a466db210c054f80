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


def test_device_pipeline_epoch_trains_without_host_batches():
    """The CPU twin of tests/test_gpu_trainer.py::test_device_pipeline_epoch_trains_without_host_batches: positives
    permuted and negatives drawn by the sampler kernel -> [K, 3, B] blocks -> persistent launches with the SGD update fused
    into the scatter (trainer.train_epoch_device), all through the emulator."""
    from fake_data import FakeDataset, base_config
    from recbole_cdr_b200.data import DeviceDomainData
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler
    from recbole_cdr_b200.utils import ModelType, get_model, get_trainer
    ds = FakeDataset(41, 60, 50, 1, 90, 80)
    rng = np.random.RandomState(0)
    su, si = ds.valid_ids('source')
    s_u, s_i = rng.choice(su, 700), rng.choice(si, 700)
    with emu_util.patched_ops(sms=2):
        smp = CrossDomainSourceSampler('train', ds, user_ids=s_u, item_ids=s_i, device='cpu').set_phase('train')
        data = DeviceDomainData(s_u, s_i, smp, device='cpu')
        blocks = list(data.epoch_blocks(64, 4, pairwise=True, generator=torch.Generator().manual_seed(1)))
        assert sum(b[0].shape[0] for b in blocks) == 700 // 64 and blocks[0][0].shape == (4, 3, 64)
        used = set(zip(s_u.tolist(), s_i.tolist()))
        valid_items = set(si.tolist())
        for u, n in zip(blocks[0][0][:, 0].reshape(-1).tolist(), blocks[0][0][:, 2].reshape(-1).tolist()):
            assert n in valid_items and (u, n) not in used
        cfg = base_config(device='cpu', latent_factor_model='BPR', source_embedding_size=64, target_embedding_size=64,
                          reg_weight=0.0, mapping_function='non_linear', mlp_hidden_size=[128], learner='sgd',
                          learning_rate=20.0, weight_decay=0.0, train_modes=['SOURCE'], epoch_num=['1'], source_split=False)
        torch.manual_seed(2022)
        model = get_model('EMCDR')(cfg, ds)
        model.set_phase('SOURCE')
        trainer = get_trainer(ModelType.CROSSDOMAIN, 'EMCDR')(cfg, model)
        losses = [trainer.train_epoch_device(data, 64, steps_per_launch=4) for _ in range(4)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0] - 1e-3, losses


def test_device_pipeline_with_row_sparse_adagrad_equals_dense_torch_adagrad():
    """train_epoch_device with ``xdr_row_optimizer: adagrad``: per batch one persistent-kernel launch (gradient rows into
    the gradient tables) + one optimizer kernel per table.  The tables after an epoch equal the oracle loss stepped by the
    dense torch.optim.Adagrad on the very same batches."""
    from fake_data import FakeDataset, base_config
    from recbole_cdr_b200.data import DeviceDomainData
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler
    from recbole_cdr_b200.utils import ModelType, get_model, get_trainer
    ds = FakeDataset(41, 60, 50, 1, 90, 80)
    rng = np.random.RandomState(0)
    su, si = ds.valid_ids('source')
    s_u, s_i = rng.choice(su, 300), rng.choice(si, 300)
    with emu_util.patched_ops(sms=2):
        smp = CrossDomainSourceSampler('train', ds, user_ids=s_u, item_ids=s_i, device='cpu').set_phase('train')
        data = DeviceDomainData(s_u, s_i, smp, device='cpu')
        cfg = base_config(device='cpu', latent_factor_model='BPR', source_embedding_size=64, target_embedding_size=64,
                          reg_weight=0.01, mapping_function='non_linear', mlp_hidden_size=[128], learner='adagrad',
                          learning_rate=0.05, weight_decay=0.0, train_modes=['SOURCE'], epoch_num=['1'], source_split=False,
                          xdr_row_optimizer='adagrad')
        torch.manual_seed(2022)
        model = get_model('EMCDR')(cfg, ds)
        model.set_phase('SOURCE')
        u0 = model.source_user_embedding.weight.detach().clone()
        i0 = model.source_item_embedding.weight.detach().clone()
        trainer = get_trainer(ModelType.CROSSDOMAIN, 'EMCDR')(cfg, model)
        # the same blocks the trainer will see: same sampler call count, same permutation seed
        smp_ref = CrossDomainSourceSampler('train', ds, user_ids=s_u, item_ids=s_i, device='cpu').set_phase('train')
        blocks = list(DeviceDomainData(s_u, s_i, smp_ref, device='cpu').epoch_blocks(
            64, 2, pairwise=True, generator=torch.Generator().manual_seed(3)))
        loss = trainer.train_epoch_device(data, 64, steps_per_launch=2, generator=torch.Generator().manual_seed(3))
    a, b = u0.clone().requires_grad_(True), i0.clone().requires_grad_(True)
    opt = torch.optim.Adagrad([a, b], lr=0.05)
    ref_total = 0.0
    for ids, _ in blocks:
        for k in range(ids.shape[0]):
            opt.zero_grad()
            l = O.emcdr_bpr_loss(a, b, ids[k, 0], ids[k, 1], ids[k, 2], 0.01)
            l.sum().backward()
            opt.step()
            ref_total += float(l.detach())
    assert abs(loss - ref_total) <= 1e-4 * abs(ref_total)
    torch.testing.assert_close(model.source_user_embedding.weight.detach(), a.detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(model.source_item_embedding.weight.detach(), b.detach(), rtol=1e-4, atol=1e-6)
    assert not model.source_user_embedding.weight.grad.any() and not model.source_item_embedding.weight.grad.any()


def test_cmf_both_mode_epoch_on_the_device_pipeline():
    """CMF BOTH mode without host batches: target batches set the epoch length, source batches restart; two persistent
    launches (source term, target term) per block with the SGD update fused.  The loss must fall over epochs."""
    from fake_data import FakeDataset, base_config
    from recbole_cdr_b200.data import DeviceDomainData
    from recbole_cdr_b200.model.cross_domain_recommender.cmf import CMF
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler, TargetDomainSampler
    from recbole_cdr_b200.trainer import CrossDomainTrainer
    ds = FakeDataset(1, 80, 70, 31, 60, 50)
    rng = np.random.RandomState(0)
    su, si = ds.valid_ids('source')
    tu, ti = ds.valid_ids('target')
    s_u, s_i = rng.choice(su, 200), rng.choice(si, 200)        # fewer source interactions: the source side restarts
    t_u, t_i = rng.choice(tu, 520), rng.choice(ti, 520)
    cfg = base_config(device='cpu', embedding_size=64, alpha=0.4, gamma=0.0, learning_rate=5.0, weight_decay=0.0,
                      learner='sgd', train_modes=['BOTH'], epoch_num=['1'])
    cfg['lambda'] = 0.0
    with emu_util.patched_ops(sms=2):
        s_smp = CrossDomainSourceSampler('train', ds, user_ids=s_u, item_ids=s_i, device='cpu').set_phase('train')
        t_smp = TargetDomainSampler(ds.num_total_user, ds.target_domain_dataset.num('target_item_id'), t_u, t_i, device='cpu')
        sd, td = DeviceDomainData(s_u, s_i, s_smp, device='cpu'), DeviceDomainData(t_u, t_i, t_smp, device='cpu')
        torch.manual_seed(1)
        m = CMF(cfg, ds)
        t = CrossDomainTrainer(cfg, m)
        w0 = m.user_embedding.weight.detach().clone()
        losses = [t.train_epoch_device_both(sd, td, 64, steps_per_launch=3) for _ in range(3)]
    steps = 520 // 32                                  # pointwise: 32 positives + 32 negatives per step
    assert np.isfinite(losses).all() and losses[-1] < losses[0] - 1e-3, losses
    assert abs(losses[0] / steps - np.log(2)) < 0.05   # alpha*BCE_s + (1-alpha)*BCE_t starts at ln 2 per step
    assert not torch.equal(w0, m.user_embedding.weight.detach())


def test_device_pipeline_pointwise_blocks_feed_the_persistent_kernel():
    """EMCDR-MF (pointwise) through train_epoch_device: the label rows share the step stride of the id rows."""
    from fake_data import FakeDataset, base_config
    from recbole_cdr_b200.data import DeviceDomainData
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler
    from recbole_cdr_b200.utils import ModelType, get_model, get_trainer
    ds = FakeDataset(41, 60, 50, 1, 90, 80)
    rng = np.random.RandomState(0)
    su, si = ds.valid_ids('source')
    s_u, s_i = rng.choice(su, 400), rng.choice(si, 400)
    with emu_util.patched_ops(sms=2):
        smp = CrossDomainSourceSampler('train', ds, user_ids=s_u, item_ids=s_i, device='cpu').set_phase('train')
        data = DeviceDomainData(s_u, s_i, smp, device='cpu')
        ids, lab = next(iter(data.epoch_blocks(64, 3, pairwise=False)))
        assert ids.shape == (3, 2, 64) and lab.shape == (3, 64) and lab.stride(0) == ids[:, 0].stride(0)
        assert lab[:, :32].all() and not lab[:, 32:].any()
        cfg = base_config(device='cpu', latent_factor_model='MF', source_embedding_size=64, target_embedding_size=64,
                          reg_weight=0.0, mapping_function='non_linear', mlp_hidden_size=[128], learner='sgd',
                          learning_rate=5.0, weight_decay=0.0, train_modes=['SOURCE'], epoch_num=['1'], source_split=False)
        torch.manual_seed(2022)
        model = get_model('EMCDR')(cfg, ds)
        model.set_phase('SOURCE')
        trainer = get_trainer(ModelType.CROSSDOMAIN, 'EMCDR')(cfg, model)
        losses = [trainer.train_epoch_device(data, 64, steps_per_launch=3) for _ in range(4)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0] - 1e-3, losses


def test_row_sharded_ranks_running_concurrently():
    """The same as above with the two ranks' persistent launches RESIDENT TOGETHER (emulator launch group): their CTAs
    interleave while they gather from and RED into each other's shards -- what happens on two GPUs."""
    from recbole_cdr_b200 import shard
    from recbole_cdr_b200.shard import RowShardedTable, train_steps_sharded
    world, K, B, dim, nu, ni = 2, 2, 64, 64, 301, 403
    ut, it, u, ip, ineg, _ = setup(nu, ni, dim, K * world, B, 13)
    L = emu_util.lib()
    with emu_util.patched_ops(sms=2, seed=5):
        def shards(full):
            tabs = [RowShardedTable.from_full(full, r, world, 'cpu') for r in range(world)]
            for t in tabs:
                t._ptrs = [s.local.data_ptr() for s in tabs]
            return tabs
        tu, ti = shards(ut), shards(it)
        gu, gi = shards(torch.zeros_like(ut)), shards(torch.zeros_like(it))
        keep, outs = [], []
        L.emu_group_begin()
        for r in range(world):
            sl = slice(r * K, (r + 1) * K)
            args = (u[sl].contiguous(), ip[sl].contiguous(), ineg[sl].contiguous())
            outs.append(train_steps_sharded(tu[r], ti[r], gu[r], gi[r], *args, reg_weight=0.01))   # queued, not run yet
            keep.append((args, dict(shard._steps_ws)))
            shard._steps_ws.clear()                   # every rank gets its own step workspace, as on its own GPU
        L.emu_group_run()
    full = lambda tabs, n: torch.stack([t.local for t in tabs], 1).reshape(-1, dim)[:n]
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    gu_ref, gi_ref = torch.zeros_like(ut), torch.zeros_like(it)
    losses = torch.cat([o[:, 0] for o in outs])
    for k in range(K * world):
        ref = O.emcdr_bpr_loss(a, b, u[k], ip[k], ineg[k], 0.01)
        torch.testing.assert_close(losses[k], ref.detach()[0], rtol=1e-4, atol=0)
        du, di = O.grads_of(ref, [a, b])
        gu_ref += du
        gi_ref += di
    torch.testing.assert_close(full(gu, nu), gu_ref, rtol=1e-4, atol=1e-4 * gu_ref.abs().max().item())
    torch.testing.assert_close(full(gi, ni), gi_ref, rtol=1e-4, atol=1e-4 * gi_ref.abs().max().item())


@pytest.mark.parametrize('K,B,dim,sms,pairwise', [(12, 96, 64, 3, True), (20, 64, 32, 2, True), (11, 64, 64, 2, False)])
def test_early_scatter_variant_for_reg_weight_zero(K, B, dim, sms, pairwise):
    """train_steps_staged_kernel<..., EARLY>: with reg_weight == 0 the scatterers do not wait for the norm exchange (opt-in
    through xdr_steps_set_early_scatter).  K > the 8-deep id / partial / norm rings, so the slot-release order is exercised.
    Per-step losses and accumulated gradients equal the oracle (and the flag changes nothing when reg_weight != 0)."""
    from recbole_cdr_b200 import _lib
    nu, ni = 300, 400
    ut, it, u, ip, ineg, y = setup(nu, ni, dim, K, B, 9, 0.3)
    L = emu_util.lib()
    results = {}
    for early in (0, 1):
        for seed in ((0,) if not early else (0, 23)):
            with emu_util.patched_ops(sms=sms, seed=seed) as ops:
                L.xdr_steps_set_early_scatter(early)
                try:
                    if pairwise:
                        out8, gu, gi = ops.train_steps(ut.clone(), it.clone(), u, ip, ineg, reg_weight=0.0)
                    else:
                        out8, gu, gi = ops.train_steps(ut.clone(), it.clone(), u, ip, None, y, loss_kind=_lib.LOSS_BCE_SIGMOID,
                                                       reg_weight=0.0)
                finally:
                    L.xdr_steps_set_early_scatter(0)
            results[(early, seed)] = (out8[:, 0].clone(), gu, gi)
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    gu_ref, gi_ref = torch.zeros_like(ut), torch.zeros_like(it)
    losses = []
    for k in range(K):
        if pairwise:
            ref = O.emcdr_bpr_loss(a, b, u[k], ip[k], ineg[k], 0.0)
        else:
            ref = O.bce_loss(torch.sigmoid(O.dot_score(a, b, u[k], ip[k])), y[k])
        losses.append(ref.detach().reshape(-1)[0])
        du, di = O.grads_of(ref, [a, b])
        gu_ref += du
        gi_ref += di
    for key, (loss, gu, gi) in results.items():
        torch.testing.assert_close(loss, torch.stack(losses), rtol=1e-4, atol=0, msg=lambda s: f'{key}: {s}')
        torch.testing.assert_close(gu, gu_ref, rtol=1e-4, atol=1e-4 * gu_ref.abs().max().item())
        torch.testing.assert_close(gi, gi_ref, rtol=1e-4, atol=1e-4 * gi_ref.abs().max().item())
    assert torch.equal(results[(0, 0)][0], results[(1, 0)][0])      # the loss path is untouched: same bits


def test_early_scatter_flag_is_ignored_when_the_norms_matter():
    K, B, dim = 5, 64, 64
    ut, it, u, ip, ineg, _ = setup(200, 250, dim, K, B, 4)
    L = emu_util.lib()
    outs = []
    for early in (0, 1):
        with emu_util.patched_ops(sms=2) as ops:
            L.xdr_steps_set_early_scatter(early)
            try:
                outs.append(ops.train_steps(ut.clone(), it.clone(), u, ip, ineg, reg_weight=0.05))
            finally:
                L.xdr_steps_set_early_scatter(0)
    for x, y in zip(outs[0], outs[1]):
        torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-8)


# ------------------------------------------------------------------------------------------ lazily zeroed gradient tables
def _oracle_grads(ut, it, u, ip, ineg, steps, reg=0.01):
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    gu, gi, losses = torch.zeros_like(ut), torch.zeros_like(it), []
    for k in steps:
        ref = O.emcdr_bpr_loss(a, b, u[k], ip[k], ineg[k], reg)
        losses.append(ref.detach()[0])
        du, di = O.grads_of(ref, [a, b])
        gu += du
        gi += di
    return gu, gi, torch.stack(losses)


@pytest.mark.parametrize('K,B,dim,sms,seed', [(6, 96, 64, 3, 0), (11, 64, 64, 2, 3), (5, 40, 32, 4, 1), (9, 96, 64, 3, 29)])
def test_lazy_gradient_tables_fresh_launch(K, B, dim, sms, seed):
    """xdr_train_steps_lazy with clear_map: the destination tables start out as GARBAGE; after the launch the rows the touch
    map marks are exactly the rows the batches name and hold exactly the oracle's gradient (the first touch zero-filled them,
    no matter which CTA or step got there first); unmarked rows were never written."""
    nu, ni = 150, 180     # small tables: many rows are named by several steps and by several CTAs
    ut, it, u, ip, ineg, _ = setup(nu, ni, dim, K, B, 5 + seed)
    junk_u, junk_i = torch.full_like(ut, 7.0), torch.full_like(it, -3.0)
    with emu_util.patched_ops(sms=sms, seed=seed) as ops:
        tm = ops.TouchMap(nu, ni, 'cpu')
        tm.words.fill_(-1)    # stale marks from an earlier optimizer step: fresh=True must clear them
        out8, gu, gi = ops.train_steps(ut.clone(), it.clone(), u, ip, ineg, reg_weight=0.01, user_dst=junk_u.clone(),
                                       item_dst=junk_i.clone(), touch=tm, fresh=True)
        su, si = tm.state()
    gu_ref, gi_ref, loss_ref = _oracle_grads(ut, it, u, ip, ineg, range(K))
    torch.testing.assert_close(out8[:, 0], loss_ref, rtol=1e-4, atol=0)
    named_u = torch.zeros(nu, dtype=torch.bool)
    named_u[u.reshape(-1)] = True
    named_i = torch.zeros(ni, dtype=torch.bool)
    named_i[ip.reshape(-1)] = True
    named_i[ineg.reshape(-1)] = True
    assert torch.equal(su != 0, named_u) and torch.equal(si != 0, named_i)
    assert bool(((su == 0) | (su == 3)).all()) and bool(((si == 0) | (si == 3)).all())   # claimed rows are all filled
    torch.testing.assert_close(gu[named_u], gu_ref[named_u], rtol=1e-4, atol=1e-4 * gu_ref.abs().max().item())
    torch.testing.assert_close(gi[named_i], gi_ref[named_i], rtol=1e-4, atol=1e-4 * gi_ref.abs().max().item())
    assert bool((gu[~named_u] == 7.0).all()) and bool((gi[~named_i] == -3.0).all())


def test_lazy_gradient_tables_accumulate_then_fresh():
    """clear_map = 0 keeps adding into the rows an earlier launch marked (and zero-fills rows it did not); a later launch with
    clear_map = 1 starts over: its marked rows hold only its own gradient."""
    K, B, dim, nu, ni = 8, 64, 64, 120, 140
    ut, it, u, ip, ineg, _ = setup(nu, ni, dim, K, B, 41)
    with emu_util.patched_ops(sms=2, seed=7) as ops:
        tm = ops.TouchMap(nu, ni, 'cpu')
        gu, gi = torch.full_like(ut, 5.0), torch.full_like(it, 5.0)
        ops.train_steps(ut.clone(), it.clone(), u[:3], ip[:3], ineg[:3], reg_weight=0.01, user_dst=gu, item_dst=gi, touch=tm,
                        fresh=True)
        ops.train_steps(ut.clone(), it.clone(), u[3:6], ip[3:6], ineg[3:6], reg_weight=0.01, user_dst=gu, item_dst=gi,
                        touch=tm, fresh=False)
        tu, ti = tm.touched()
        gu_ref, gi_ref, _ = _oracle_grads(ut, it, u, ip, ineg, range(6))
        torch.testing.assert_close(gu[tu], gu_ref[tu], rtol=1e-4, atol=1e-4 * gu_ref.abs().max().item())
        torch.testing.assert_close(gi[ti], gi_ref[ti], rtol=1e-4, atol=1e-4 * gi_ref.abs().max().item())
        assert bool((gu[~tu] == 5.0).all()) and bool((gi[~ti] == 5.0).all())
        ops.train_steps(ut.clone(), it.clone(), u[6:], ip[6:], ineg[6:], reg_weight=0.01, user_dst=gu, item_dst=gi, touch=tm,
                        fresh=True)
        tu2, ti2 = tm.touched()
    gu_ref2, gi_ref2, _ = _oracle_grads(ut, it, u, ip, ineg, range(6, K))
    named = torch.zeros(nu, dtype=torch.bool)
    named[u[6:].reshape(-1)] = True
    assert torch.equal(tu2, named)
    torch.testing.assert_close(gu[tu2], gu_ref2[tu2], rtol=1e-4, atol=1e-4 * gu_ref2.abs().max().item())
    torch.testing.assert_close(gi[ti2], gi_ref2[ti2], rtol=1e-4, atol=1e-4 * gi_ref2.abs().max().item())


def test_lazy_gradient_tables_pointwise_and_refusals():
    from recbole_cdr_b200 import _lib
    K, B, dim, nu, ni = 4, 64, 64, 90, 110
    ut, it, u, i, _, y = setup(nu, ni, dim, K, B, 13, 0.3)
    with emu_util.patched_ops(sms=2, seed=2) as ops:
        tm = ops.TouchMap(nu, ni, 'cpu')
        out8, gu, gi = ops.train_steps(ut.clone(), it.clone(), u, i, None, y, loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=0.0,
                                       user_dst=torch.full_like(ut, 9.0), item_dst=torch.full_like(it, 9.0), touch=tm, fresh=True)
        tu, ti = tm.touched()
        # the weight tables themselves cannot be lazily zeroed destinations
        w_u, w_i = ut.clone(), it.clone()
        with pytest.raises(_lib.XdrError, match='cannot be the weight table'):
            ops.train_steps(w_u, w_i, u, i, None, y, loss_kind=_lib.LOSS_BCE_SIGMOID, user_dst=w_u, item_dst=w_i, touch=tm)
    a, b = ut.clone().requires_grad_(True), it.clone().requires_grad_(True)
    gu_ref, gi_ref = torch.zeros_like(ut), torch.zeros_like(it)
    for s in range(K):
        ref = O.bce_loss(torch.sigmoid(O.dot_score(a, b, u[s], i[s])), y[s])
        du, di = O.grads_of(ref, [a, b])
        gu_ref += du
        gi_ref += di
    torch.testing.assert_close(gu[tu], gu_ref[tu], rtol=1e-4, atol=1e-4 * gu_ref.abs().max().item())
    torch.testing.assert_close(gi[ti], gi_ref[ti], rtol=1e-4, atol=1e-4 * gi_ref.abs().max().item())
    assert bool((gu[~tu] == 9.0).all())


# ------------------------------------------------------------------------------------------ hot rows in shared memory
@pytest.mark.parametrize('pairwise,sms,seed', [(True, 3, 0), (True, 2, 9), (False, 2, 4)])
def test_hot_rows_accumulate_in_shared_memory_and_flush_once(pairwise, sms, seed):
    """xdr_steps_set_hot_rows: the gradients of the listed (popular) rows are added in every CTA's shared memory over the whole
    launch and flushed once; everything else goes the usual way.  Same sums as the oracle, for lists that name popular rows,
    rows nobody touches, an out-of-range id and a duplicate."""
    from recbole_cdr_b200 import _lib
    K, B, dim, nu, ni = 7, 96, 64, 120, 150
    ut, it, u, ip, ineg, y = setup(nu, ni, dim, K, B, 21 + seed)
    ip[:, ::3] = 5          # a third of every batch names item 5, a sixth item 9 (the popular rows)
    ineg[:, ::6] = 9
    u[:, ::4] = 2
    hot_u = torch.tensor([2, 77, 2, 10_000], dtype=torch.int64)       # popular, cold, duplicate, out of range
    hot_i = torch.tensor([5, 9, 149, 33], dtype=torch.int64)
    with emu_util.patched_ops(sms=sms, seed=seed) as ops:
        ops.set_steps_hot_rows(hot_u, hot_i)
        try:
            if pairwise:
                out8, gu, gi = ops.train_steps(ut.clone(), it.clone(), u, ip, ineg, reg_weight=0.01)
            else:
                out8, gu, gi = ops.train_steps(ut.clone(), it.clone(), u, ip, None, y, loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=0.01)
        finally:
            ops.set_steps_hot_rows(None, None)
        if pairwise:
            o2, gu2, gi2 = ops.train_steps(ut.clone(), it.clone(), u, ip, ineg, reg_weight=0.01)
        else:
            o2, gu2, gi2 = ops.train_steps(ut.clone(), it.clone(), u, ip, None, y, loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=0.01)
    assert torch.equal(out8[:, 0], o2[:, 0])
    torch.testing.assert_close(gu, gu2, rtol=1e-5, atol=1e-6 * gu2.abs().max().item())
    torch.testing.assert_close(gi, gi2, rtol=1e-5, atol=1e-6 * gi2.abs().max().item())
    if pairwise:
        gu_ref, gi_ref, _ = _oracle_grads(ut, it, u, ip, ineg, range(K))
        torch.testing.assert_close(gu, gu_ref, rtol=1e-4, atol=1e-4 * gu_ref.abs().max().item())
        torch.testing.assert_close(gi, gi_ref, rtol=1e-4, atol=1e-4 * gi_ref.abs().max().item())
