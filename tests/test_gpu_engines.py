"""GPU parity tests of the second-generation kernels: tensor-core row-tile MLPs (mma.sync and tcgen05), the fused CoNet tower
pass, row-sparse optimizers, fused score + top-k (both engines), early scatter, the five further models.  All of them were
written in round 1 without GPU access (logic under the CPU CTA emulator, tests/test_emu_*.py) and ran on a B200 for the first
time in round 2's first GPU call (profiles/r2_call1_summary.txt): every test of this file passed there except the ones
whose docstrings say what was changed since."""
import numpy as np
import pytest
import torch

from fake_data import FakeDataset, base_config  # noqa: F401
from golden_util import Golden
from oracle import cdr_oracle as O
from test_gpu_kernels import LOSS_RTOL, dev, lib, ops, rand_ids, rand_table
from test_gpu_models import EMCDR_CFG, build, check_loss_and_grads, cuda_batch

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------ tensor-core fused MLP (tc_mlp.cu)
@pytest.mark.parametrize('batch', [1, 31, 32, 33, 63, 64, 65, 1000, 8192, 20000])
def test_tc_mlp_map_loss_matches_oracle(batch):
    """EMCDR map step through the 3xTF32 tile kernel at sizes around the 32- and 64-row tiles."""
    g = torch.Generator().manual_seed(111)
    src, tgt = rand_table(3000, 64, 112, 0.3), rand_table(3000, 64, 113, 0.3)
    ws = [torch.randn(128, 64, generator=g) * 0.2, torch.randn(64, 128, generator=g) * 0.2]
    bs = [torch.randn(128, generator=g) * 0.1, torch.randn(64, generator=g) * 0.1]
    idx = rand_ids(batch, 3000, 114, 1.3)
    leaves = [t.clone().requires_grad_(True) for t in [src, tgt] + ws + bs]
    ref = O.emcdr_map_loss(leaves[0], leaves[1], idx.view(-1, 1), leaves[2:4], leaves[4:6])
    ref.backward()
    c = [t.to(dev()).requires_grad_(True) for t in [src, tgt] + ws + bs]
    assert ops().fused_mlp_supported([64, 128, 64], 'tc')
    loss = ops().fused_mlp_loss(0, 0, lib().ACT_TANH, idx.to(dev()), None, None, (c[0], None, None, None, c[1]), c[2:4], c[4:6],
                                'tc')
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=LOSS_RTOL, atol=0)
    (loss * 1.3).backward()
    for got, want, nm in zip(c, leaves, ('src', 'tgt', 'W1', 'W2', 'b1', 'b2')):
        atol = max(1e-7, 1e-4 * want.grad.abs().max().item())
        torch.testing.assert_close(got.grad.cpu(), want.grad * 1.3, rtol=2e-4, atol=atol, msg=lambda s: f'{nm}: {s}')


# ------------------------------------------------------------------- early-scatter form of the persistent kernel (reg_weight 0)
@pytest.mark.parametrize('pairwise', [True, False])
def test_train_steps_early_scatter_matches_the_default_kernel(pairwise):
    """reg_weight == 0: the scatterers of train_steps_staged_kernel<..., EARLY> do not wait for the norm exchange.  Same losses
    (bit for bit: the loss path is untouched) and the same gradient tables as the hardware-validated default kernel."""
    K, B, dim, nu, ni = 40, 8192, 64, 200_000, 300_000
    g = torch.Generator().manual_seed(5)
    ut, it = (torch.randn(nu, dim, generator=g) * 0.1).to(dev()), (torch.randn(ni, dim, generator=g) * 0.1).to(dev())
    u = torch.randint(0, nu, (K, B), generator=g).to(dev())
    ia, ib = torch.randint(0, ni, (K, B), generator=g).to(dev()), torch.randint(0, ni, (K, B), generator=g).to(dev())
    y = (torch.rand(K, B, generator=g) < 0.5).float().to(dev())
    outs = []
    for early in (False, True):
        ops().set_steps_early_scatter(early)
        try:
            if pairwise:
                outs.append(ops().train_steps(ut, it, u, ia, ib, reg_weight=0.0))
            else:
                outs.append(ops().train_steps(ut, it, u, ia, None, y, loss_kind=lib().LOSS_BCE_SIGMOID, reg_weight=0.0))
        finally:
            ops().set_steps_early_scatter(False)
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0][:, 0], outs[1][0][:, 0])
    for a, b in ((outs[0][1], outs[1][1]), (outs[0][2], outs[1][2])):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6 * float(a.abs().max()))


# ---------------------------------------------------------------------------------- tcgen05 fused map step (tc5_mlp.cu)
@pytest.mark.parametrize('batch', [1, 127, 128, 129, 1000, 8192, 40000])
def test_tc5_mlp_map_loss_matches_oracle(batch):
    """EMCDR map step with all six products of a 128-row tile on tcgen05.mma (bf16x3, TMEM-resident weight gradients);
    40000 rows give every CTA more than one tile (accumulation in tensor memory over tiles)."""
    g = torch.Generator().manual_seed(211)
    src, tgt = rand_table(3000, 64, 212, 0.3), rand_table(3000, 64, 213, 0.3)
    ws = [torch.randn(128, 64, generator=g) * 0.2, torch.randn(64, 128, generator=g) * 0.2]
    bs = [torch.randn(128, generator=g) * 0.1, torch.randn(64, generator=g) * 0.1]
    idx = rand_ids(batch, 3000, 214, 1.3)
    leaves = [t.clone().requires_grad_(True) for t in [src, tgt] + ws + bs]
    ref = O.emcdr_map_loss(leaves[0], leaves[1], idx.view(-1, 1), leaves[2:4], leaves[4:6])
    ref.backward()
    c = [t.to(dev()).requires_grad_(True) for t in [src, tgt] + ws + bs]
    assert ops().fused_mlp_supported([64, 128, 64], 'tc5') and not ops().fused_mlp_supported([128, 32, 16, 1], 'tc5')
    loss = ops().fused_mlp_loss(0, 0, lib().ACT_TANH, idx.to(dev()), None, None, (c[0], None, None, None, c[1]), c[2:4], c[4:6],
                                'tc5')
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=LOSS_RTOL, atol=0)
    (loss * 1.3).backward()
    for got, want, nm in zip(c, leaves, ('src', 'tgt', 'W1', 'W2', 'b1', 'b2')):
        atol = max(1e-7, 1e-4 * want.grad.abs().max().item())
        torch.testing.assert_close(got.grad.cpu(), want.grad * 1.3, rtol=2e-4, atol=atol, msg=lambda s: f'{nm}: {s}')


@pytest.mark.parametrize('case', ['non_linear', 'items'])
def test_emcdr_map_phase_tc5_engine(case):
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden(f'emcdr_map_{case}')
    m = build(EMCDR, g, dict(EMCDR_CFG, latent_factor_model='BPR', mapping_function='non_linear', xdr_fused_mlp='tc5'))
    assert m.fused_mlp_engine == 'tc5'
    m.set_phase('OVERLAP')
    check_loss_and_grads(m, g, cuda_batch(g), grad_rtol=2e-4, grad_atol=2e-6)


@pytest.mark.parametrize('batch,factor', [(8192, 1.0), (1000, 1.3), (129, 0.25)])
def test_tc5_mlp_eager_step_in_place_gradients(batch, factor):
    """'inplace' table gradients + the tcgen05 engine: forward accumulates every gradient in ONE launch (upstream gradient 1),
    backward() runs the correction pass ((g - 1)-fold; over at once for g = 1).  Same loss and gradients as the oracle."""
    g = torch.Generator().manual_seed(311)
    src, tgt = rand_table(3000, 64, 312, 0.3), rand_table(3000, 64, 313, 0.3)
    ws = [torch.randn(128, 64, generator=g) * 0.2, torch.randn(64, 128, generator=g) * 0.2]
    bs = [torch.randn(128, generator=g) * 0.1, torch.randn(64, generator=g) * 0.1]
    idx = rand_ids(batch, 3000, 314, 1.3)
    leaves = [t.clone().requires_grad_(True) for t in [src, tgt] + ws + bs]
    ref = O.emcdr_map_loss(leaves[0], leaves[1], idx.view(-1, 1), leaves[2:4], leaves[4:6])
    (ref * factor).backward()
    c = [t.to(dev()).requires_grad_(True) for t in [src, tgt] + ws + bs]
    prev = ops().get_table_grad_mode()
    ops().set_table_grad_mode('inplace')
    seen, real = [], ops().call
    ops().call = lambda nm, *a, **k: (seen.append(a[21]) if nm == 'xdr_tc5_mlp_step' else None, real(nm, *a, **k))[1]
    try:
        loss = ops().fused_mlp_loss(0, 0, lib().ACT_TANH, idx.to(dev()), None, None, (c[0], None, None, None, c[1]), c[2:4], c[4:6],
                                    'tc5')
        (loss * factor).backward()
    finally:
        ops().call = real
        ops().set_table_grad_mode(prev)
    assert seen == [1, 2]
    torch.testing.assert_close(loss.detach().cpu(), ref.detach(), rtol=LOSS_RTOL, atol=0)
    for got, want, nm in zip(c, leaves, ('src', 'tgt', 'W1', 'W2', 'b1', 'b2')):
        atol = max(1e-7, 1e-4 * want.grad.abs().max().item())
        torch.testing.assert_close(got.grad.cpu(), want.grad, rtol=2e-4, atol=atol, msg=lambda s: f'{nm}: {s}')


def test_tc_mlp_supported_stacks():
    assert ops().fused_mlp_supported([64, 128, 64], 'tc') and ops().fused_mlp_supported([128, 32, 16, 1], 'tc')
    assert not ops().fused_mlp_supported([512, 64, 1], 'tc')
    assert not ops().fused_mlp_supported([64, 12, 64], 'tc')


@pytest.mark.parametrize('case,mf', [('non_linear', 'non_linear'), ('linear', 'linear'), ('items', 'non_linear')])
def test_emcdr_map_phase_tc_engine(case, mf):
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden(f'emcdr_map_{case}')
    m = build(EMCDR, g, dict(EMCDR_CFG, latent_factor_model='BPR', mapping_function=mf, xdr_fused_mlp='tc'))
    assert m.fused_mlp_engine == 'tc'
    m.set_phase('OVERLAP')
    check_loss_and_grads(m, g, cuda_batch(g), grad_atol=1e-6)


def test_dtcdr_tc_engine():
    from recbole_cdr_b200.model.cross_domain_recommender.dtcdr import DTCDR
    g = Golden('dtcdr_neumf')
    m = build(DTCDR, g, dict(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF',
                             alpha=g.meta('alpha'), xdr_fused_mlp='tc'))
    assert m._fused_ok() and m.fused_mlp_engine == 'tc'
    batch = cuda_batch(g)
    check_loss_and_grads(m, g, batch, grad_rtol=2e-4, grad_atol=2e-6)
    torch.testing.assert_close(m.predict(batch).cpu(), g.t('predict'), rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------------------------------ fused CoNet tower pass (tc_conet.cu)
@pytest.mark.parametrize('tag', ['users', 'items'])
def test_conet_fused_tower_kernel(tag):
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    g = Golden(f'conet_{tag}')
    m = build(CoNet, g, dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8], xdr_fused_conet=True))
    assert m._fused_ok()
    check_loss_and_grads(m, g, cuda_batch(g), grad_rtol=2e-4, grad_atol=2e-6)


def _relu_knife_edges(tabs, P, user, item, n_ov, eps=1e-5):
    """[B] bool: interactions with some cross-stitch pre-activation |z| < eps (fp64 forward of conet.py:105-142)."""
    import torch.nn.functional as F
    d = torch.float64
    x_s = torch.cat([tabs['source_user'][user], tabs['source_item'][item]], 1).to(d)
    x_t = torch.cat([tabs['target_user'][user], tabs['target_item'][item]], 1).to(d)
    m = (user < n_ov).to(d).unsqueeze(1)
    edge = torch.zeros(user.shape[0], dtype=torch.bool)
    for l in range(len(P['ws'])):
        cross = P['h'][l].to(d).t()
        zs = F.linear(x_s, P['ws'][l].to(d), P['bs'][l].to(d)) + m * (x_t @ cross)
        zt = F.linear(x_t, P['wt'][l].to(d), P['bt'][l].to(d)) + m * (x_s @ cross)
        edge |= (torch.cat([zs, zt], 1).abs() < eps).any(1)
        x_s, x_t = torch.relu(zs), torch.relu(zt)
    return edge


@pytest.mark.parametrize('batch,dim,hidden,want', [(1, 32, [16], 0), (63, 32, [32, 16, 8], 1), (1000, 64, [64, 32, 16, 8], 0),
                                                   (16384, 128, [64, 32, 16, 8], 1)])
def test_conet_fused_matches_oracle(batch, dim, hidden, want):
    """One tower pass (loss, probabilities via the loss, every gradient) against the oracle at sizes up to config #3's."""
    gen = torch.Generator().manual_seed(5)
    n_u, n_i, n_ov = 5000, 3000, 2500
    names = ('source_user', 'source_item', 'target_user', 'target_item')
    tabs = {k: rand_table(n_u if 'user' in k else n_i, dim, 300 + j, 0.3) for j, k in enumerate(names)}
    dims = [2 * dim] + list(hidden)
    mk = lambda a, b, s=0.25: torch.randn(b, a, generator=gen) * s
    P = dict(ws=[mk(a, b) for a, b in zip(dims[:-1], dims[1:])], wt=[mk(a, b) for a, b in zip(dims[:-1], dims[1:])],
             h=[mk(a, b) for a, b in zip(dims[:-1], dims[1:])],
             bs=[torch.randn(b, generator=gen) * 0.1 for b in dims[1:]], bt=[torch.randn(b, generator=gen) * 0.1 for b in dims[1:]],
             out_s_w=mk(dims[-1], 1, 0.5), out_s_b=torch.randn(1, generator=gen) * 0.1,
             out_t_w=mk(dims[-1], 1, 0.5), out_t_b=torch.randn(1, generator=gen) * 0.1)
    user, item = rand_ids(batch, n_u, 7, 1.3), rand_ids(batch, n_i, 8)
    # ReLU is not differentiable at 0: an interaction with a pre-activation within rounding distance of 0 gets a different
    # (equally valid) sub-gradient from any implementation whose dot products round differently.  Round 2, call 2: with
    # 16384 rows one interaction (position 2938, layer-2 pre-activation 2.9e-7 in fp64) made the 3xTF32 kernel differ from the
    # fp32 oracle by exactly that interaction's gradient (scripts/diag_conet_hot.py; no race: compute-sanitizer racecheck clean,
    # bit-reproducible, the CPU emulator agrees with the oracle).  Such interactions get another item.
    for _ in range(8):
        edge = _relu_knife_edges(tabs, P, user, item, n_ov)
        if not bool(edge.any()):
            break
        item = torch.where(edge, (item + 1) % n_i, item)
    label = (torch.rand(batch, generator=gen) < 0.5).float()
    lt = {k: v.clone().requires_grad_(True) for k, v in tabs.items()}
    lp = {k: ([x.clone().requires_grad_(True) for x in v] if isinstance(v, list) else v.clone().requires_grad_(True))
          for k, v in P.items()}
    ps, pt = O.conet_towers(lt, user, item, lp, True, n_ov)
    ref = O.bce_loss(ps if want == 0 else pt, label)
    ref.backward()
    ct = {k: v.to(dev()).requires_grad_(True) for k, v in tabs.items()}
    cp = {k: ([x.to(dev()).requires_grad_(True) for x in v] if isinstance(v, list) else v.to(dev()).requires_grad_(True))
          for k, v in P.items()}
    sfx = 's' if want == 0 else 't'
    assert ops().conet_fused_supported(dims, dim)
    loss = ops().conet_tower_loss(want, False, n_ov, user.to(dev()), item.to(dev()), label.to(dev()),
                                  tuple(ct[k] for k in names), cp[f'out_{sfx}_w'], cp[f'out_{sfx}_b'], cp['ws'], cp['bs'],
                                  cp['wt'], cp['bt'], cp['h'])
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=LOSS_RTOL, atol=0)
    loss.backward()

    def chk(got, want_t, nm):
        w = torch.zeros_like(got).cpu() if want_t.grad is None else want_t.grad
        atol = max(1e-7, 1e-4 * w.abs().max().item())
        torch.testing.assert_close(got.grad.cpu() if got.grad is not None else torch.zeros_like(w), w, rtol=2e-4, atol=atol,
                                   msg=lambda s: f'{nm}: {s}')

    for k in names:
        chk(ct[k], lt[k], k)
    for key in ('ws', 'bs', 'wt', 'bt', 'h'):
        for l in range(len(hidden)):
            chk(cp[key][l], lp[key][l], f'{key}[{l}]')
    chk(cp[f'out_{sfx}_w'], lp[f'out_{sfx}_w'], 'out_w')
    chk(cp[f'out_{sfx}_b'], lp[f'out_{sfx}_b'], 'out_b')


# ------------------------------------------------------------------------------------------ row-sparse optimizers (sparse_optim.cu)
@pytest.mark.parametrize('kind,name', [(0, 'sgd'), (1, 'adagrad'), (2, 'adam')])
def test_sparse_optim_rows_matches_oracle(kind, name):
    from oracle import optim_oracle as OO
    rng = np.random.RandomState(3)
    n, d, b = 5000, 64, 4096
    w0 = (rng.randn(n, d) * 0.1).astype(np.float32)
    w = torch.from_numpy(w0.copy()).to(dev())
    stamp = torch.zeros(n, dtype=torch.int32, device=dev())
    s1, s2 = torch.zeros_like(w), torch.zeros_like(w)
    rw, rs, rm, rv = w0.copy(), np.zeros_like(w0), np.zeros_like(w0), np.zeros_like(w0)
    for step in range(1, 4):
        ids = np.minimum(rng.zipf(1.3, b) - 1, n - 1)
        rows = (rng.randn(b, d) * 0.05).astype(np.float32)
        g = np.zeros_like(w0)
        np.add.at(g, ids, rows)
        gd = torch.from_numpy(g.copy()).to(dev())
        ops().sparse_optim_rows(kind, w, gd, torch.from_numpy(ids).to(dev()), stamp, step, 0.05,
                                state1=s1 if kind else None, state2=s2 if kind == 2 else None, adam_t=step,
                                eps=1e-10 if kind == 1 else 1e-8)
        assert not gd.any().item(), 'gradient rows must be zero after the step'
        if kind == 0:
            rw = OO.sgd_step(rw, g, 0.05)
        elif kind == 1:
            rw, rs = OO.adagrad_step(rw, rs, g, 0.05)
        else:
            rw, rm, rv = OO.sparse_adam_step(rw, rm, rv, g, ids, step, 0.05)
        np.testing.assert_allclose(w.cpu().numpy(), rw, rtol=1e-5, atol=1e-7, err_msg=name)


def test_trainer_row_sparse_adagrad_follows_dense_torch_adagrad():
    """CMF, three batches: tables stepped by the row-sparse kernel == tables stepped by dense torch.optim.Adagrad."""
    from recbole_cdr_b200.data import Interaction
    from recbole_cdr_b200.model.cross_domain_recommender.cmf import CMF
    from recbole_cdr_b200.trainer import CrossDomainTrainer
    ds = FakeDataset(1, 300, 280, 120, 200, 150)
    rng = np.random.RandomState(0)
    su, si = ds.valid_ids('source')
    tu, ti = ds.valid_ids('target')
    batches = []
    for _ in range(3):
        batches.append(Interaction({
            'source_user_id': torch.from_numpy(rng.choice(su, 512)), 'source_item_id': torch.from_numpy(rng.choice(si, 512)),
            'source_label': torch.from_numpy((rng.rand(512) < 0.5).astype(np.float32)),
            'target_user_id': torch.from_numpy(rng.choice(tu, 512)), 'target_item_id': torch.from_numpy(rng.choice(ti, 512)),
            'target_label': torch.from_numpy((rng.rand(512) < 0.5).astype(np.float32))}))
    cfg = dict(embedding_size=64, alpha=0.3, gamma=0.1, learning_rate=0.05, weight_decay=0.0, train_modes=['BOTH'],
               epoch_num=['1'], learner='adagrad')
    cfg['lambda'] = 0.1
    models = []
    for row_opt in (None, 'adagrad'):
        torch.manual_seed(7)
        c = base_config(**cfg)
        if row_opt:
            c['xdr_row_optimizer'] = row_opt
        m = CMF(c, ds).to('cuda')
        t = CrossDomainTrainer(c, m)
        t._train_epoch(batches, 0)
        models.append(m)
    for (n1, p1), (_, p2) in zip(models[0].named_parameters(), models[1].named_parameters()):
        torch.testing.assert_close(p2, p1, rtol=1e-4, atol=1e-6, msg=lambda s: f'{n1}: {s}')


# ------------------------------------------------------------------------------------------ fused full-sort top-k (topk_score.cu)
@pytest.mark.parametrize('engine', ['mma', 'tc5'])
@pytest.mark.parametrize('B,n_items,D,k', [(5, 300, 64, 10), (200, 50000, 64, 20), (64, 1000, 128, 100), (4096, 200001, 64, 10)])
def test_full_sort_topk_matches_masked_torch_topk(B, n_items, D, k, engine):
    if engine == 'tc5' and D > 64:
        pytest.skip('the tcgen05 engine takes dim <= 64')
    rng = np.random.RandomState(B)
    U = torch.from_numpy((rng.randn(B, D) * 0.3).astype(np.float32))
    I = torch.from_numpy((rng.randn(n_items, D) * 0.3).astype(np.float32))
    lens = rng.randint(0, 30, B)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ids = np.concatenate([np.sort(rng.choice(np.arange(1, n_items), l, replace=False)) for l in lens] + [np.zeros(0, np.int64)])
    hp, hi = torch.from_numpy(ptr).to(dev()), torch.from_numpy(ids.astype(np.int64)).to(dev())
    sc, pos = ops().full_sort_topk(U.to(dev()), I.to(dev()), k, hist_ptr=hp, hist_ids=hi, engine=engine)
    full = (U.to(dev()) @ I.to(dev()).T)
    full[:, 0] = -float('inf')
    rows = torch.repeat_interleave(torch.arange(B, device=dev()), torch.from_numpy(lens).to(dev()))
    full[rows, hi] = -float('inf')
    rs, ri = torch.topk(full, k, dim=1)
    torch.testing.assert_close(sc, rs, rtol=2e-5, atol=1e-6)
    # every returned id carries its score, is unique, and is not masked; ids agree wherever scores are well separated
    torch.testing.assert_close(torch.gather(full, 1, pos), sc, rtol=2e-5, atol=1e-6)
    assert (torch.sort(pos, dim=1).values.diff(dim=1) > 0).all()
    gap = (rs[:, :-1] - rs[:, 1:]) > 1e-4 * rs[:, :-1].abs().clamp_min(1e-3)
    sep = torch.cat([torch.ones_like(gap[:, :1]), gap], 1) & torch.cat([gap, torch.ones_like(gap[:, :1])], 1)
    assert torch.equal(pos[sep], ri[sep])


# ------------------------------------------------------------------------------------------ F4: the five remaining drop-in models
def build_f4(model_cls, g, cfg, with_edges=False):
    from fake_data import FakeDatasetF4
    ds = (FakeDatasetF4 if with_edges else FakeDataset).from_golden(g)
    torch.manual_seed(0)
    m = model_cls(base_config(**cfg), ds)
    state = {n: g.param(n) for n in g.param_names()}
    for k in [k for k in state if k.startswith('seq.')]:   # DeepAPF's second name of its item MLP (deepapf.py:56-62)
        state['item_mlp.' + k[len('seq.'):]] = state[k]
    m.load_state_dict(state, strict=True)
    return m.to('cuda')


def check_full_sort_gpu(m, g):
    """full_sort_predict (and the fused full_sort_topk where the model has one) against the reference's scores."""
    import variants_util as V
    if not g.has('full_sort_predict'):
        return
    with torch.no_grad():
        got = m.full_sort_predict(V.batch(g, 'cuda', 'fbatch/'))
    torch.testing.assert_close(got.cpu().reshape(-1), g.t('full_sort_predict').reshape(-1), rtol=1e-4, atol=2e-6)
    if hasattr(m, 'full_sort_topk'):
        V.check_topk_against_reference(m, g, 'cuda')


def test_f4_clfm():
    from recbole_cdr_b200.model.cross_domain_recommender.clfm import CLFM
    g = Golden('f4_clfm')
    m = build_f4(CLFM, g, dict(user_embedding_size=64, source_item_embedding_size=64, target_item_embedding_size=64,
                               share_embedding_size=32, alpha=g.meta('alpha'), reg_weight=g.meta('reg_weight')))
    batch = cuda_batch(g)
    check_loss_and_grads(m, g, batch, grad_rtol=2e-4, grad_atol=2e-6)
    torch.testing.assert_close(m.predict(batch).cpu(), g.t('predict'), rtol=1e-5, atol=1e-6)
    check_full_sort_gpu(m, g)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_f4_deepapf(tag):
    from recbole_cdr_b200.model.cross_domain_recommender.deepapf import DeepAPF
    g = Golden(f'f4_deepapf_{tag}')
    m = build_f4(DeepAPF, g, dict(embedding_size=64, beta=0.5))
    batch = cuda_batch(g)
    check_loss_and_grads(m, g, batch, grad_rtol=2e-4, grad_atol=2e-6)
    torch.testing.assert_close(m.predict(batch).cpu(), g.t('predict'), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('case,phase', [('source', 'SOURCE'), ('target', 'TARGET'), ('map_users', 'OVERLAP'), ('map_items', 'OVERLAP')])
def test_f4_sscdr(case, phase):
    from recbole_cdr_b200.model.cross_domain_recommender.sscdr import SSCDR
    g = Golden(f'f4_sscdr_{case}')
    m = build_f4(SSCDR, g, {'embedding_size': 64, 'margin': 1, 'mlp_hidden_size': [128], 'lambda': 0.25}, with_edges=True)
    m.set_phase(phase)
    if g.has('meta/np_seed'):
        np.random.seed(g.meta('np_seed'))
    check_loss_and_grads(m, g, cuda_batch(g), grad_rtol=2e-4, grad_atol=2e-6)
    check_full_sort_gpu(m, g)


@pytest.mark.parametrize('tag', ['items', 'users'])
@pytest.mark.parametrize('phase', ['source', 'target'])
def test_f4_natr(tag, phase):
    from recbole_cdr_b200.model.cross_domain_recommender.natr import NATR
    g = Golden(f'f4_natr_{tag}_{phase}')
    m = build_f4(NATR, g, dict(source_embedding_size=64, target_embedding_size=64, reg_weight=1e-3,
                               max_inter_length=g.meta('max_inter_length')), with_edges=True)
    m.set_phase(phase.upper())
    batch = cuda_batch(g)
    check_loss_and_grads(m, g, batch, grad_rtol=2e-4, grad_atol=2e-6)
    torch.testing.assert_close(m.predict(batch).cpu(), g.t('predict'), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_f4_dcdcsr_four_stages(tag):
    from recbole_cdr_b200.model.cross_domain_recommender.dcdcsr import DCDCSR
    cfg = dict(latent_factor_model='BPR', embedding_size=64, mlp_hidden_size=[128], k=5, map_batch_size=64)
    g = Golden(f'f4_dcdcsr_{tag}_source1')
    m = build_f4(DCDCSR, g, cfg, with_edges=True)
    m.set_phase('SOURCE')
    check_loss_and_grads(m, g, cuda_batch(g))
    g = Golden(f'f4_dcdcsr_{tag}_target1')
    m.set_phase('TARGET')
    check_loss_and_grads(m, g, cuda_batch(g))
    check_full_sort_gpu(m, g)
    g = Golden(f'f4_dcdcsr_{tag}_both')
    m.set_phase('BOTH')
    torch.testing.assert_close(m.benchmark_embedding.cpu(), g.t('benchmark_embedding'), rtol=1e-4, atol=1e-6)
    np.random.seed(g.meta('np_seed'))
    check_loss_and_grads(m, g, cuda_batch(g), grad_atol=1e-6)
    g = Golden(f'f4_dcdcsr_{tag}_target2')
    m.set_phase('TARGET')
    torch.testing.assert_close(m.affine_embedding.cpu(), g.t('affine_embedding'), rtol=1e-4, atol=1e-6)
    check_loss_and_grads(m, g, cuda_batch(g))
    check_full_sort_gpu(m, g)


def test_device_pipeline_with_row_sparse_adagrad_equals_dense_torch_adagrad():
    """GPU twin of tests/test_emu_steps.py: train_epoch_device + xdr_row_optimizer == dense torch Adagrad on the oracle loss."""
    from recbole_cdr_b200.data import DeviceDomainData
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler
    from recbole_cdr_b200.utils import ModelType, get_model, get_trainer
    ds = FakeDataset(201, 300, 280, 1, 500, 450)
    rng = np.random.RandomState(0)
    su, si = ds.valid_ids('source')
    s_u, s_i = rng.choice(su, 6000), rng.choice(si, 6000)
    cfg = base_config(latent_factor_model='BPR', source_embedding_size=64, target_embedding_size=64, reg_weight=0.01,
                      mapping_function='non_linear', mlp_hidden_size=[128], learner='adagrad', learning_rate=0.05,
                      weight_decay=0.0, train_modes=['SOURCE'], epoch_num=['1'], source_split=False, xdr_row_optimizer='adagrad')
    torch.manual_seed(2022)
    model = get_model('EMCDR')(cfg, ds).to('cuda')
    model.set_phase('SOURCE')
    u0 = model.source_user_embedding.weight.detach().cpu().clone()
    i0 = model.source_item_embedding.weight.detach().cpu().clone()
    trainer = get_trainer(ModelType.CROSSDOMAIN, 'EMCDR')(cfg, model)
    mk = lambda: DeviceDomainData(s_u, s_i, CrossDomainSourceSampler('train', ds, user_ids=s_u, item_ids=s_i,
                                                                     device='cuda').set_phase('train'))
    blocks = list(mk().epoch_blocks(1024, 2, pairwise=True, generator=torch.Generator(device='cuda').manual_seed(3)))
    loss = trainer.train_epoch_device(mk(), 1024, steps_per_launch=2, generator=torch.Generator(device='cuda').manual_seed(3))
    a, b = u0.clone().requires_grad_(True), i0.clone().requires_grad_(True)
    opt = torch.optim.Adagrad([a, b], lr=0.05)
    ref_total = 0.0
    for ids, _ in blocks:
        ids = ids.cpu()
        for k in range(ids.shape[0]):
            opt.zero_grad()
            l = O.emcdr_bpr_loss(a, b, ids[k, 0], ids[k, 1], ids[k, 2], 0.01)
            l.sum().backward()
            opt.step()
            ref_total += float(l.detach())
    assert abs(loss - ref_total) <= 1e-4 * abs(ref_total)
    # Adagrad's FIRST update of an element is lr * g / (|g| + 1e-10): +-lr whatever |g| is -- unless |g| is within a few
    # orders of magnitude of 1e-10, where the ratio amplifies the last bits of g (the scatter-add sums a row's duplicates in
    # another order than index_add).  Round 2, call 1: 27 of 49 984 elements differed by up to 1.3e-4 = lr * 2.6e-3 for that
    # reason.  So: every element within lr * 1e-2, and all but a handful at the usual tolerance.
    for got, want in ((model.source_user_embedding.weight.detach().cpu(), a.detach()),
                      (model.source_item_embedding.weight.detach().cpu(), b.detach())):
        err = (got - want).abs()
        assert float(err.max()) <= 0.05 * 1e-2
        assert float((err > 1e-6 + 1e-4 * want.abs()).float().mean()) <= 2e-3


# ------------------------------------------------------------------------------------------ tcgen05 building blocks (tc5.cuh)
def test_tc5_selftest_gemm_tf32_k_major():
    """D = A B^T on tcgen05 (3xTF32, TMEM accumulator) with the operands staged K-major by the kernel itself: the hardware
    check of the descriptor reading, through the library (scripts/ubench_tcgen05.cu is the standalone form)."""
    g = torch.Generator().manual_seed(0)
    N, K = 64, 64
    A, B = torch.randn(128, K, generator=g).to(dev()), torch.randn(N, K, generator=g).to(dev())
    D = torch.zeros(128, N, device=dev())
    lib().call('xdr_tc5_selftest', A.data_ptr(), B.data_ptr(), N, K, 0, 0, D.data_ptr(), lib().cur_stream())
    torch.cuda.synchronize()
    torch.testing.assert_close(D.cpu().double(), A.cpu().double() @ B.cpu().double().T, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('a_mn,b_mn', [(0, 1), (1, 0), (1, 1)])
def test_tc5_selftest_refuses_mn_major_tf32(a_mn, b_mn):
    """Round 2, call 1: with SWIZZLE_NONE neither stride-field assignment reproduces the product for MN-major 32-bit operands
    on a B200 (profiles/r2_ubench_tcgen05.txt).  No kernel of the library uses them; the self-test now refuses the request."""
    A, B, D = torch.zeros(128, 64, device=dev()), torch.zeros(64, 64, device=dev()), torch.zeros(128, 64, device=dev())
    with pytest.raises(lib().XdrError, match='MN-major TF32'):
        lib().call('xdr_tc5_selftest', A.data_ptr(), B.data_ptr(), 64, 64, a_mn, b_mn, D.data_ptr(), lib().cur_stream())


@pytest.mark.parametrize('a_mn,b_mn', [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)])
def test_tc5_selftest_gemm_bf16x3_all_majors(a_mn, b_mn):
    """The same on kind::f16 with bf16 hi / lo planes (bf16x3): the operand format planned for the tcgen05 training kernels."""
    g = torch.Generator().manual_seed(10 + a_mn * 2 + b_mn)
    N, K = 64, 64
    A, B = torch.randn(128, K, generator=g).to(dev()), torch.randn(N, K, generator=g).to(dev())
    D = torch.zeros(128, N, device=dev())
    lib().call('xdr_tc5_selftest_bf16', A.data_ptr(), B.data_ptr(), N, K, a_mn, b_mn, D.data_ptr(), lib().cur_stream())
    torch.cuda.synchronize()
    ref = A.cpu().double() @ B.cpu().double().T
    mass = A.cpu().double().abs() @ B.cpu().double().abs().T
    assert bool(((D.cpu().double() - ref).abs() <= 2.0 ** -15 * mass + 1e-6).all())


# ---------------------------------------------------------------------------------------------------------------------------
# the second set of reference goldens (tests/golden/v_*.npz) through the kernels that have not met hardware yet
# ---------------------------------------------------------------------------------------------------------------------------
import variants_util as _V  # noqa: E402


@pytest.mark.parametrize('engine', ['fma', 'tc'])
@pytest.mark.parametrize('name', [n for n in _V.VARIANTS if _V.spec(Golden(n))['model'] in ('EMCDR', 'DTCDR')])
def test_variant_fused_mlp_engines(name, engine):
    g = Golden(name)
    _V.check_against_reference(_V.build(g, 'cuda', xdr_fused_mlp=engine), g, 'cuda')


@pytest.mark.parametrize('name', [n for n in _V.VARIANTS if _V.spec(Golden(n))['model'] == 'CoNet'])
def test_variant_fused_conet(name):
    g = Golden(name)
    _V.check_against_reference(_V.build(g, 'cuda', xdr_fused_conet=True), g, 'cuda')


@pytest.mark.parametrize('name', [n for n in _V.VARIANTS
                                  if Golden(n).has('full_sort_predict') and _V.spec(Golden(n))['model'] in ('EMCDR', 'CMF', 'BiTGCF')])
def test_variant_fused_topk_matches_the_reference_scores(name):
    g = Golden(name)
    _V.check_topk_against_reference(_V.build(g, 'cuda'), g, 'cuda')
