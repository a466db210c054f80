"""The drop-in model classes on CPU tensors THROUGH the CTA emulator: ``recbole_cdr_b200.ops`` is patched
(``emu_util.patched_ops``) so that the fused-kernel entry points run the real kernel sources under tests/emu, and the
results are compared with tests/golden/*.npz -- outputs of the UNMODIFIED reference classes.  This exercises the Python
glue (argument order of the ctypes calls, gradient routing of the autograd functions, state_dict keys) and the kernel logic
of the paths written without GPU access; it does not replace the hardware parity tests."""
import pytest
import torch

import emu_util
from fake_data import FakeDataset, base_config
from golden_util import Golden

LOSS_RTOL = 1e-4


def build_cpu(model_cls, g, cfg):
    ds = FakeDataset.from_golden(g, None)
    torch.manual_seed(0)
    m = model_cls(base_config(device='cpu', **cfg), ds)
    m.load_state_dict({n: g.param(n) for n in g.param_names()}, strict=True)
    return m


def cpu_batch(g, prefix='batch/'):
    from recbole_cdr_b200.data import Interaction
    return Interaction({k[len(prefix):]: torch.from_numpy(g.z[k]) for k in g.z.files if k.startswith(prefix)})


def check(m, g, batch, grad_rtol=2e-4, grad_atol=2e-6):
    m.zero_grad()
    loss = m.calculate_loss(batch)
    losses = list(loss) if isinstance(loss, tuple) else [loss]
    for got, ref in zip(losses, g.losses()):
        torch.testing.assert_close(got.detach().reshape(-1), ref.reshape(-1), rtol=LOSS_RTOL, atol=0)
    sum(l.sum() for l in losses).backward()
    for name, p in m.named_parameters():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        torch.testing.assert_close(got, g.grad(name), rtol=grad_rtol, atol=grad_atol, msg=lambda s: f'grad {name}: {s}')


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_conet_fused_kernel_vs_reference_golden(tag):
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    g = Golden(f'conet_{tag}')
    with emu_util.patched_ops():
        m = build_cpu(CoNet, g, dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8], xdr_fused_conet=True))
        assert m._fused_ok()
        check(m, g, cpu_batch(g))


@pytest.mark.parametrize('engine', ['fma', 'tc'])
@pytest.mark.parametrize('case,mf', [('non_linear', 'non_linear'), ('linear', 'linear'), ('items', 'non_linear')])
def test_emcdr_map_phase_vs_reference_golden(case, mf, engine):
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden(f'emcdr_map_{case}')
    cfg = dict(source_embedding_size=64, target_embedding_size=64, reg_weight=0.01, mlp_hidden_size=[128],
               latent_factor_model='BPR', mapping_function=mf, xdr_fused_mlp=engine)
    with emu_util.patched_ops():
        m = build_cpu(EMCDR, g, cfg)
        assert m.fused_mlp_engine == engine
        m.set_phase('OVERLAP')
        check(m, g, cpu_batch(g), grad_atol=1e-6)


@pytest.mark.parametrize('engine', ['fma', 'tc'])
def test_dtcdr_vs_reference_golden(engine):
    from recbole_cdr_b200.model.cross_domain_recommender.dtcdr import DTCDR
    g = Golden('dtcdr_neumf')
    with emu_util.patched_ops():
        m = build_cpu(DTCDR, g, dict(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF',
                                     alpha=g.meta('alpha'), xdr_fused_mlp=engine))
        assert m._fused_ok() and m.fused_mlp_engine == engine
        batch = cpu_batch(g)
        check(m, g, batch)
        torch.testing.assert_close(m.predict(batch), g.t('predict'), rtol=1e-4, atol=1e-6)


def test_unpatched_ops_still_refuse_cpu_tensors():
    """The patch is scoped: outside the context manager the product path has no CPU route."""
    from recbole_cdr_b200 import ops
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.gather_rows_raw(torch.zeros(8, 64), torch.zeros(2, dtype=torch.int64))


# ---------------------------------------------------------------------------------------------------------------------------
# The hardware-validated composed paths through the emulator: pins the emulator itself (the same kernels are parity-green
# on a B200, tests/test_gpu_models.py) and keeps these paths checkable on CPU after every change.
# ---------------------------------------------------------------------------------------------------------------------------
EMCDR_CFG = dict(source_embedding_size=64, target_embedding_size=64, reg_weight=0.01, mlp_hidden_size=[128])


@pytest.mark.parametrize('lfm', ['bpr', 'mf'])
@pytest.mark.parametrize('phase', ['source', 'target'])
def test_emcdr_rec_phases_composed(lfm, phase):
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden(f'emcdr_{lfm}_{phase}')
    with emu_util.patched_ops():
        m = build_cpu(EMCDR, g, dict(EMCDR_CFG, latent_factor_model=lfm.upper(), mapping_function='non_linear'))
        m.set_phase(phase.upper())
        batch = cpu_batch(g)
        check(m, g, batch, grad_rtol=1e-4, grad_atol=2e-7)
        torch.testing.assert_close(m.predict(batch), g.t('predict'), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('case,mf', [('non_linear', 'non_linear'), ('linear', 'linear'), ('items', 'non_linear')])
def test_emcdr_map_phase_composed(case, mf):
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden(f'emcdr_map_{case}')
    with emu_util.patched_ops():
        m = build_cpu(EMCDR, g, dict(EMCDR_CFG, latent_factor_model='BPR', mapping_function=mf, xdr_fused_mlp=False))
        m.set_phase('OVERLAP')
        check(m, g, cpu_batch(g), grad_rtol=1e-4, grad_atol=1e-6)
        pred = m.predict(cpu_batch(g, 'pbatch/'))
        torch.testing.assert_close(pred, g.t('predict_overlap_phase'), rtol=1e-4, atol=1e-6)


def test_cmf_composed():
    from recbole_cdr_b200.model.cross_domain_recommender.cmf import CMF
    g = Golden('cmf_both')
    with emu_util.patched_ops():
        m = build_cpu(CMF, g, {'embedding_size': 64, 'alpha': g.meta('alpha'), 'lambda': g.meta('lambda'),
                               'gamma': g.meta('gamma')})
        batch = cpu_batch(g)
        check(m, g, batch, grad_rtol=1e-4, grad_atol=2e-7)
        torch.testing.assert_close(m.predict(batch), g.t('predict'), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_conet_composed(tag):
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    g = Golden(f'conet_{tag}')
    with emu_util.patched_ops():
        m = build_cpu(CoNet, g, dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8]))
        batch = cpu_batch(g)
        check(m, g, batch)
        torch.testing.assert_close(m.predict(batch), g.t('predict'), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_conet_two_pass_form_still_matches_the_golden(tag):
    """``xdr_stack_passes: False`` keeps one tower pass per domain batch (the form the stacked pass replaced)."""
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    g = Golden(f'conet_{tag}')
    with emu_util.patched_ops():
        m = build_cpu(CoNet, g, dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8], xdr_stack_passes=False))
        assert not m.stack_passes
        check(m, g, cpu_batch(g))


@pytest.mark.parametrize('n_s,n_t', [(37, 64), (64, 5), (1, 1)])
def test_conet_stacked_pass_equals_two_passes_on_ragged_halves(n_s, n_t):
    """The two halves of a BOTH batch may differ in length on the last batch (dataloader.py:148-162): the stacked pass splits
    at the right row (odd row counts: the second half's views start at unaligned row offsets)."""
    from recbole_cdr_b200.data import Interaction
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    g = Golden('conet_users')
    b = cpu_batch(g)
    gen = torch.Generator().manual_seed(n_s * 100 + n_t)
    pick = lambda t, n: t[torch.randint(0, t.numel(), (n,), generator=gen)]
    batch = {}
    for dom, n in (('source', n_s), ('target', n_t)):
        for k in b.columns:
            if k.startswith(dom):
                batch[k] = pick(b[k], n)
    batch = Interaction(batch)
    res = []
    with emu_util.patched_ops():
        for stacked in (True, False):
            m = build_cpu(CoNet, g, dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8],
                                         xdr_stack_passes=stacked))
            m.zero_grad()
            loss = m.calculate_loss(batch)
            loss.backward()
            res.append((loss.detach().clone(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}))
    torch.testing.assert_close(res[0][0], res[1][0], rtol=1e-6, atol=0)
    assert res[0][1].keys() == res[1][1].keys()
    for n in res[0][1]:
        torch.testing.assert_close(res[0][1][n], res[1][1][n], rtol=1e-5, atol=1e-7, msg=lambda s: f'grad {n}: {s}')


def test_dtcdr_composed():
    from recbole_cdr_b200.model.cross_domain_recommender.dtcdr import DTCDR
    g = Golden('dtcdr_neumf')
    with emu_util.patched_ops():
        m = build_cpu(DTCDR, g, dict(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF',
                                     alpha=g.meta('alpha'), xdr_fused_mlp=False))
        batch = cpu_batch(g)
        check(m, g, batch)
        torch.testing.assert_close(m.predict(batch), g.t('predict'), rtol=1e-4, atol=1e-6)


def test_trainer_row_sparse_adagrad_follows_dense_torch_adagrad():
    """CMF, three batches: tables stepped by the row-sparse kernel == tables stepped by dense torch.optim.Adagrad
    (trainer._train_epoch_row_sparse; the CPU twin of the test in tests/test_gpu_engines.py)."""
    import numpy as np
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
            'source_user_id': torch.from_numpy(rng.choice(su, 128)), 'source_item_id': torch.from_numpy(rng.choice(si, 128)),
            'source_label': torch.from_numpy((rng.rand(128) < 0.5).astype(np.float32)),
            'target_user_id': torch.from_numpy(rng.choice(tu, 128)), 'target_item_id': torch.from_numpy(rng.choice(ti, 128)),
            'target_label': torch.from_numpy((rng.rand(128) < 0.5).astype(np.float32))}))
    cfg = dict(embedding_size=64, alpha=0.3, gamma=0.1, learning_rate=0.05, weight_decay=0.0, train_modes=['BOTH'],
               epoch_num=['1'], learner='adagrad', device='cpu')
    cfg['lambda'] = 0.1
    models = []
    with emu_util.patched_ops():
        for row_opt in (None, 'adagrad'):
            torch.manual_seed(7)
            c = base_config(**cfg)
            if row_opt:
                c['xdr_row_optimizer'] = row_opt
            m = CMF(c, ds)
            t = CrossDomainTrainer(c, m)
            t._train_epoch(batches, 0)
            models.append(m)
    for (n1, p1), (_, p2) in zip(models[0].named_parameters(), models[1].named_parameters()):
        torch.testing.assert_close(p2, p1, rtol=1e-4, atol=1e-6, msg=lambda s: f'{n1}: {s}')


def test_full_sort_topk_agrees_with_full_sort_predict():
    """EMCDR (OVERLAP phase: mapped user vectors) and CMF: the fused top-k equals masking + topk of full_sort_predict."""
    import numpy as np
    from recbole_cdr_b200.data import Interaction
    from recbole_cdr_b200.model.cross_domain_recommender.cmf import CMF
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    with emu_util.patched_ops():
        g = Golden('emcdr_map_non_linear')
        m = build_cpu(EMCDR, g, dict(EMCDR_CFG, latent_factor_model='BPR', mapping_function='non_linear'))
        m.set_phase('OVERLAP')
        g2 = Golden('cmf_both')
        m2 = build_cpu(CMF, g2, {'embedding_size': 64, 'alpha': g2.meta('alpha'), 'lambda': g2.meta('lambda'),
                                 'gamma': g2.meta('gamma')})
        for model in (m, m2):
            users = torch.arange(1, 12)
            inter = Interaction({'target_user_id': users})
            full = model.full_sort_predict(inter).view(len(users), -1)
            rng = np.random.RandomState(0)
            ptr, ids = [0], []
            for _ in users:
                h = np.unique(rng.randint(1, full.shape[1], 9))
                ids.append(h)
                ptr.append(ptr[-1] + len(h))
            hp, hi = torch.tensor(ptr), torch.from_numpy(np.concatenate(ids))
            sc, pos = model.full_sort_topk(inter, 10, hp, hi)
            ref = full.clone()
            ref[:, 0] = -float('inf')
            for u in range(len(users)):
                ref[u, hi[hp[u]:hp[u + 1]]] = -float('inf')
            rs, ri = torch.topk(ref, 10, dim=1)
            torch.testing.assert_close(sc, rs, rtol=2e-5, atol=1e-6)
            assert torch.equal(pos, ri)


@pytest.mark.parametrize('way', ['concat', 'mean'])
def test_bitgcf_composed(way):
    """BiTGCF (graph SpMM, propagate, transfer + normalise kernels: graph_prop.cu, hardware-validated) through the emulator."""
    from golden_util import bitgcf_graph
    from recbole_cdr_b200.model.cross_domain_recommender.bitgcf import BiTGCF
    g = Golden(f'bitgcf_{way}')
    _, _, edges, _ = bitgcf_graph(g)
    with emu_util.patched_ops():
        ds = FakeDataset.from_golden(g, edges)
        m = BiTGCF(base_config(device='cpu', embedding_size=32, n_layers=2, reg_weight=0.001, lambda_source=0.8,
                               lambda_target=0.7, drop_rate=0.0, connect_way=way), ds)
        m.load_state_dict({n: g.param(n) for n in g.param_names()}, strict=True)
        batch = cpu_batch(g)
        check(m, g, batch, grad_rtol=2e-4, grad_atol=2e-7)
        torch.testing.assert_close(m.predict(batch), g.t('predict'), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('name,golden,cfg,phase', [
    ('EMCDR', 'emcdr_bpr_source', dict(EMCDR_CFG, latent_factor_model='BPR', mapping_function='non_linear'), 'SOURCE'),
    ('EMCDR', 'emcdr_mf_target', dict(EMCDR_CFG, latent_factor_model='MF', mapping_function='non_linear'), 'TARGET'),
    ('EMCDR', 'emcdr_map_non_linear', dict(EMCDR_CFG, latent_factor_model='BPR', mapping_function='non_linear'), 'OVERLAP'),
    ('CMF', 'cmf_both', {'embedding_size': 64, 'alpha': 0.3, 'lambda': 0.05, 'gamma': 0.02}, None),
    ('CoNet', 'conet_users', dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8]), None),
    ('DTCDR', 'dtcdr_neumf', dict(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF', alpha=0.4), None),
])
def test_touched_rows_cover_every_row_with_a_gradient(name, golden, cfg, phase):
    """The row-sparse optimizer visits only ``model.touched_rows(batch)``: every table row that receives a gradient must be
    in that list (and the list must not name rows of tables the loss does not read)."""
    import importlib
    cls = getattr(importlib.import_module(f'recbole_cdr_b200.model.cross_domain_recommender.{name.lower()}'), name)
    g = Golden(golden)
    with emu_util.patched_ops():
        m = build_cpu(cls, g, cfg)
        if phase:
            m.set_phase(phase)
        batch = cpu_batch(g)
        m.zero_grad()
        loss = m.calculate_loss(batch)
        (sum(loss) if isinstance(loss, tuple) else loss).sum().backward()
        touched = {}
        for table, ids in m.touched_rows(batch):
            touched.setdefault(id(table), set()).update(ids.reshape(-1).tolist())
        for pname, p in m.named_parameters():
            if not pname.endswith('_embedding.weight'):
                continue
            rows = set(torch.nonzero(p.grad.abs().sum(dim=1)).reshape(-1).tolist()) if p.grad is not None else set()
            assert rows <= touched.get(id(p), set()), f'{pname}: rows with a gradient that touched_rows() does not list'
