"""The five remaining drop-in models (CLFM, DeepAPF, SSCDR, NATR, DCDCSR) against
tests/golden/f4_*.npz -- outputs of the UNMODIFIED reference classes (oracle/make_golden_f4.py) -- on CPU tensors through the
CTA emulator (``emu_util.patched_ops``): same state_dict keys (strict load), same loss, every parameter gradient, predict.
The GPU counterparts live in tests/test_gpu_engines.py."""
import numpy as np
import pytest
import torch

import emu_util
from fake_data import FakeDataset, FakeDatasetF4, base_config
from golden_util import Golden
from test_emu_models import check, cpu_batch


def build(model_cls, g, cfg, with_edges=False):
    ds = (FakeDatasetF4 if with_edges else FakeDataset).from_golden(g)
    torch.manual_seed(0)
    m = model_cls(base_config(device='cpu', **cfg), ds)
    state = {n: g.param(n) for n in g.param_names()}
    # DeepAPF registers its item MLP under two names (seq / item_mlp, deepapf.py:56-62); the golden holds
    # named_parameters(), which reports each tensor once
    for k in [k for k in state if k.startswith('seq.')]:
        state['item_mlp.' + k[len('seq.'):]] = state[k]
    m.load_state_dict(state, strict=True)
    assert [n for n, _ in m.named_parameters()] == g.param_names(), 'parameter names / order differ from the reference'
    return m


def check_full_sort(m, g):
    """``full_sort_predict`` in the model's current phase against the reference's (goldens that carry one); where the model
    has a fused ``full_sort_topk``, that too (against torch.topk of the reference's masked scores)."""
    import variants_util as V
    assert g.has('full_sort_predict')
    with torch.no_grad():
        got = m.full_sort_predict(cpu_batch(g, 'fbatch/'))
    torch.testing.assert_close(got.reshape(-1), g.t('full_sort_predict').reshape(-1), rtol=1e-4, atol=2e-6)
    if hasattr(m, 'full_sort_topk'):
        V.check_topk_against_reference(m, g, 'cpu')


def test_clfm():
    from recbole_cdr_b200.model.cross_domain_recommender.clfm import CLFM
    g = Golden('f4_clfm')
    with emu_util.patched_ops():
        m = build(CLFM, g, dict(user_embedding_size=64, source_item_embedding_size=64, target_item_embedding_size=64,
                                share_embedding_size=32, alpha=g.meta('alpha'), reg_weight=g.meta('reg_weight')))
        batch = cpu_batch(g)
        check(m, g, batch)
        torch.testing.assert_close(m.predict(batch), g.t('predict'), rtol=1e-5, atol=1e-6)
        check_full_sort(m, g)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_deepapf(tag):
    from recbole_cdr_b200.model.cross_domain_recommender.deepapf import DeepAPF
    g = Golden(f'f4_deepapf_{tag}')
    with emu_util.patched_ops():
        m = build(DeepAPF, g, dict(embedding_size=64, beta=0.5))
        batch = cpu_batch(g)
        check(m, g, batch)
        torch.testing.assert_close(m.predict(batch), g.t('predict'), rtol=1e-5, atol=1e-6)


SSCDR_CFG = {'embedding_size': 64, 'margin': 1, 'mlp_hidden_size': [128], 'lambda': 0.25}


@pytest.mark.parametrize('phase', ['source', 'target'])
def test_sscdr_rec_phases(phase):
    from recbole_cdr_b200.model.cross_domain_recommender.sscdr import SSCDR
    g = Golden(f'f4_sscdr_{phase}')
    with emu_util.patched_ops():
        m = build(SSCDR, g, SSCDR_CFG, with_edges=True)
        m.set_phase(phase.upper())
        batch = cpu_batch(g)
        check(m, g, batch)
        torch.testing.assert_close(m.predict(batch), g.t('predict'), rtol=1e-5, atol=1e-6)
        if phase == 'target':
            check_full_sort(m, g)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_sscdr_map_phase(tag):
    """The positive / negative draws use NumPy's global RNG in the reference's call order: same seed, same samples."""
    from recbole_cdr_b200.model.cross_domain_recommender.sscdr import SSCDR
    g = Golden(f'f4_sscdr_map_{tag}')
    with emu_util.patched_ops():
        m = build(SSCDR, g, SSCDR_CFG, with_edges=True)
        m.set_phase('OVERLAP')
        np.random.seed(g.meta('np_seed'))
        check(m, g, cpu_batch(g), grad_atol=1e-6)
        torch.testing.assert_close(m.predict(cpu_batch(g, 'pbatch/')), g.t('predict_overlap_phase'), rtol=1e-4, atol=1e-6)
        check_full_sort(m, g)


@pytest.mark.parametrize('tag', ['items', 'users'])
@pytest.mark.parametrize('phase', ['source', 'target'])
def test_natr(tag, phase):
    from recbole_cdr_b200.model.cross_domain_recommender.natr import NATR
    g = Golden(f'f4_natr_{tag}_{phase}')
    with emu_util.patched_ops():
        m = build(NATR, g, dict(source_embedding_size=64, target_embedding_size=64, reg_weight=1e-3,
                                max_inter_length=g.meta('max_inter_length')), with_edges=True)
        m.set_phase(phase.upper())
        batch = cpu_batch(g)
        check(m, g, batch)
        torch.testing.assert_close(m.predict(batch), g.t('predict'), rtol=1e-5, atol=1e-6)


DCDCSR_CFG = dict(latent_factor_model='BPR', embedding_size=64, mlp_hidden_size=[128], k=5, map_batch_size=64)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_dcdcsr_four_stages(tag):
    """SOURCE #1, TARGET #1, BOTH (benchmark embedding + map loss), TARGET #2 (affine embedding): each stage starts from
    the reference's parameters at that stage and must reproduce its loss, gradients, predict and derived embeddings."""
    from recbole_cdr_b200.model.cross_domain_recommender.dcdcsr import DCDCSR
    with emu_util.patched_ops():
        g = Golden(f'f4_dcdcsr_{tag}_source1')
        m = build(DCDCSR, g, DCDCSR_CFG, with_edges=True)
        m.set_phase('SOURCE')
        check(m, g, cpu_batch(g))
        torch.testing.assert_close(m.predict(cpu_batch(g)), g.t('predict'), rtol=1e-5, atol=1e-6)
        g = Golden(f'f4_dcdcsr_{tag}_target1')
        m.set_phase('TARGET')
        check(m, g, cpu_batch(g))
        torch.testing.assert_close(m.predict(cpu_batch(g)), g.t('predict'), rtol=1e-5, atol=1e-6)
        check_full_sort(m, g)
        g = Golden(f'f4_dcdcsr_{tag}_both')
        m.set_phase('BOTH')
        torch.testing.assert_close(m.benchmark_embedding, g.t('benchmark_embedding'), rtol=1e-4, atol=1e-6)
        np.random.seed(g.meta('np_seed'))
        check(m, g, cpu_batch(g), grad_atol=1e-6)
        g = Golden(f'f4_dcdcsr_{tag}_target2')
        m.set_phase('TARGET')
        torch.testing.assert_close(m.affine_embedding, g.t('affine_embedding'), rtol=1e-4, atol=1e-6)
        check(m, g, cpu_batch(g))
        torch.testing.assert_close(m.predict(cpu_batch(g)), g.t('predict'), rtol=1e-5, atol=1e-6)
        check_full_sort(m, g)


def test_get_model_and_trainer_resolve_the_new_classes():
    from recbole_cdr_b200.utils import ModelType, get_model, get_trainer
    for name in ('CLFM', 'DeepAPF', 'SSCDR', 'NATR', 'DCDCSR'):
        assert get_model(name).__name__ == name
    assert get_trainer(ModelType.CROSSDOMAIN, 'DCDCSR').__name__ == 'DCDCSRTrainer'
    assert get_trainer(ModelType.CROSSDOMAIN, 'SSCDR').__name__ == 'CrossDomainTrainer'


@pytest.mark.parametrize('name,golden,cfg,phase,edges', [
    ('CLFM', 'f4_clfm', dict(user_embedding_size=64, source_item_embedding_size=64, target_item_embedding_size=64,
                             share_embedding_size=32, alpha=0.3, reg_weight=1e-2), None, False),
    ('DeepAPF', 'f4_deepapf_users', dict(embedding_size=64, beta=0.5), None, False),
    ('DeepAPF', 'f4_deepapf_items', dict(embedding_size=64, beta=0.5), None, False),
    ('SSCDR', 'f4_sscdr_source', SSCDR_CFG, 'SOURCE', True),
])
def test_touched_rows_cover_every_row_with_a_gradient(name, golden, cfg, phase, edges):
    import importlib
    cls = getattr(importlib.import_module(f'recbole_cdr_b200.model.cross_domain_recommender.{name.lower()}'), name)
    g = Golden(golden)
    with emu_util.patched_ops():
        m = build(cls, g, cfg, with_edges=edges)
        if phase:
            m.set_phase(phase)
        batch = cpu_batch(g)
        m.zero_grad()
        m.calculate_loss(batch).sum().backward()
        touched = {}
        for table, ids in m.touched_rows(batch):
            touched.setdefault(id(table), set()).update(ids.reshape(-1).tolist())
        for pname, p in m.named_parameters():
            if pname.endswith('_embedding.weight') and p.grad is not None:
                rows = set(torch.nonzero(p.grad.abs().sum(dim=1)).reshape(-1).tolist())
                assert rows <= touched.get(id(p), set()), f'{pname}: rows with a gradient that touched_rows() does not list'
