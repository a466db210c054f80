"""GPU parity of the drop-in model classes against tests/golden/*.npz -- outputs of the UNMODIFIED reference classes.

``load_state_dict(strict=True)`` from the reference's own parameter names doubles as the state_dict-key check."""
import pytest
import torch

from fake_data import FakeDataset, base_config
from golden_util import Golden

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_RTOL, GRAD_ATOL = 1e-4, 2e-7


def build(model_cls, g, cfg, edges=None):
    ds = FakeDataset.from_golden(g, edges)
    torch.manual_seed(0)
    m = model_cls(base_config(**cfg), ds)
    state = {n: g.param(n) for n in g.param_names()}
    m.load_state_dict(state, strict=True)  # same keys and shapes as the reference's state_dict
    return m.to('cuda')


def cuda_batch(g, prefix='batch/'):
    from recbole_cdr_b200.data import Interaction
    return Interaction({k[len(prefix):]: torch.from_numpy(g.z[k]) for k in g.z.files if k.startswith(prefix)}).to('cuda')


def check_loss_and_grads(m, g, batch, grad_rtol=GRAD_RTOL, grad_atol=GRAD_ATOL):
    m.zero_grad()
    loss = m.calculate_loss(batch)
    losses = list(loss) if isinstance(loss, tuple) else [loss]
    assert len(losses) == len(g.losses())
    for got, ref in zip(losses, g.losses()):
        torch.testing.assert_close(got.detach().cpu().reshape(-1), ref.reshape(-1), rtol=LOSS_RTOL, atol=0)
    sum(l.sum() for l in losses).backward()
    for name, p in m.named_parameters():
        got = p.grad.cpu() if p.grad is not None else torch.zeros_like(p).cpu()
        torch.testing.assert_close(got, g.grad(name), rtol=grad_rtol, atol=grad_atol, msg=lambda s: f'grad {name}: {s}')
    return losses


EMCDR_CFG = dict(source_embedding_size=64, target_embedding_size=64, reg_weight=0.01, mlp_hidden_size=[128])


@pytest.mark.parametrize('lfm', ['bpr', 'mf'])
@pytest.mark.parametrize('phase', ['source', 'target'])
def test_emcdr_rec_phases(lfm, phase):
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden(f'emcdr_{lfm}_{phase}')
    m = build(EMCDR, g, dict(EMCDR_CFG, latent_factor_model=lfm.upper(), mapping_function='non_linear'))
    m.set_phase(phase.upper())
    batch = cuda_batch(g)
    losses = check_loss_and_grads(m, g, batch)
    assert losses[0].shape == (1,)
    torch.testing.assert_close(m.predict(batch).cpu(), g.t('predict'), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('fused', [True, False])
@pytest.mark.parametrize('case,mf', [('non_linear', 'non_linear'), ('linear', 'linear'), ('items', 'non_linear')])
def test_emcdr_map_phase_and_predict(case, mf, fused):
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden(f'emcdr_map_{case}')
    m = build(EMCDR, g, dict(EMCDR_CFG, latent_factor_model='BPR', mapping_function=mf, xdr_fused_mlp=fused))
    assert m.mode == ('overlap_items' if case == 'items' else 'overlap_users')
    m.set_phase('OVERLAP')
    batch = cuda_batch(g)
    assert batch['overlap'].dim() == 2  # the reference's [b, 1] overlap batch
    check_loss_and_grads(m, g, batch, grad_atol=1e-6)
    pred = m.predict(cuda_batch(g, 'pbatch/'))
    torch.testing.assert_close(pred.cpu(), g.t('predict_overlap_phase'), rtol=1e-4, atol=1e-6)
    # full_sort_predict of the OVERLAP phase == predict over every target item
    pb = cuda_batch(g, 'pbatch/')
    users = pb['target_user_id'][:7]
    from recbole_cdr_b200.data import Interaction
    full = m.full_sort_predict(Interaction({'target_user_id': users})).view(7, -1)
    n_items = full.shape[1]
    assert n_items == m.target_num_items
    some_items = torch.arange(1, n_items, 5, device='cuda')
    for r in range(7):
        pr = m.predict(Interaction({'target_user_id': users[r].repeat(some_items.numel()), 'target_item_id': some_items}))
        torch.testing.assert_close(full[r, some_items], pr, rtol=1e-4, atol=1e-5)


def test_cmf():
    from recbole_cdr_b200.model.cross_domain_recommender.cmf import CMF
    g = Golden('cmf_both')
    m = build(CMF, g, {'embedding_size': 64, 'alpha': g.meta('alpha'), 'lambda': g.meta('lambda'), 'gamma': g.meta('gamma')})
    batch = cuda_batch(g)
    check_loss_and_grads(m, g, batch)
    torch.testing.assert_close(m.predict(batch).cpu(), g.t('predict'), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_conet(tag):
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    g = Golden(f'conet_{tag}')
    m = build(CoNet, g, dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8]))
    batch = cuda_batch(g)
    check_loss_and_grads(m, g, batch, grad_rtol=2e-4, grad_atol=2e-6)
    pred = m.predict(batch)
    assert pred.shape == (len(batch['target_user_id']), 1)
    torch.testing.assert_close(pred.cpu(), g.t('predict'), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_conet_two_pass_form(tag):
    """``xdr_stack_passes: False``: one tower pass per domain batch (the stacked pass of ``calculate_loss`` is the default)."""
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    g = Golden(f'conet_{tag}')
    m = build(CoNet, g, dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8], xdr_stack_passes=False))
    assert not m.stack_passes
    check_loss_and_grads(m, g, cuda_batch(g), grad_rtol=2e-4, grad_atol=2e-6)


@pytest.mark.parametrize('engine', [0, 1, 2])
@pytest.mark.parametrize('n_s,n_t', [(37, 64), (1000, 333), (1, 1), (4096, 4096)])
def test_conet_stacked_pass_equals_two_passes_on_ragged_halves(n_s, n_t, engine):
    """The halves of a BOTH batch may differ in length (dataloader.py:148-162); odd row counts make the second half's views
    start at rows that are not multiples of anything.  Every dense engine of the cross-stitch layers (``xdr_dense_engine``:
    fp32 tiles; tcgen05, whose input-gradient kernel takes the accumulating form in ``ops.cross_pair``; the per-call mix)."""
    from recbole_cdr_b200 import _lib
    from recbole_cdr_b200.data import Interaction
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    g = Golden('conet_users')
    b = cuda_batch(g)
    gen = torch.Generator().manual_seed(n_s * 100 + n_t)
    batch = {}
    for dom, n in (('source', n_s), ('target', n_t)):
        sel = None
        for k in b.columns:
            if k.startswith(dom):
                if sel is None:
                    sel = torch.randint(0, b[k].numel(), (n,), generator=gen).cuda()
                batch[k] = b[k][sel]
    batch = Interaction(batch)
    res = []
    for stacked in (True, False):
        m = build(CoNet, g, dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8], xdr_stack_passes=stacked,
                                 xdr_dense_engine=engine))
        assert m.dense_engine == engine
        m.zero_grad()
        loss = m.calculate_loss(batch)
        loss.backward()
        torch.cuda.synchronize()
        res.append((loss.detach().cpu(), {n: p.grad.cpu() for n, p in m.named_parameters() if p.grad is not None}))
    assert _lib._lib.xdr_set_dense_engine(0) == 0     # the per-call engine never leaks into the library's setting
    torch.testing.assert_close(res[0][0], res[1][0], rtol=1e-5, atol=0)
    assert res[0][1].keys() == res[1][1].keys()
    for n in res[0][1]:
        scale = max(1e-6, float(res[1][1][n].abs().max()))
        torch.testing.assert_close(res[0][1][n], res[1][1][n], rtol=1e-4, atol=2e-5 * scale, msg=lambda s: f'grad {n}: {s}')


def test_conet_default_engine_is_tcgen05_and_the_fp32_tiles_agree():
    """CoNet's cross-stitch layers run on the tcgen05 dense engine by default (bf16x3 products, fp32-faithful): against the
    golden (``test_conet``) and, here, against the same step on the fp32 FMA tiles -- different arithmetic, same result."""
    from recbole_cdr_b200.data import Interaction
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    g = Golden('conet_users')
    b = cuda_batch(g)
    gen = torch.Generator().manual_seed(7)
    batch = {}
    for dom in ('source', 'target'):
        sel = None
        for k in b.columns:
            if k.startswith(dom):
                if sel is None:
                    sel = torch.randint(0, b[k].numel(), (2048,), generator=gen).cuda()
                batch[k] = b[k][sel]
    batch = Interaction(batch)
    res = []
    for cfg in ({}, {'xdr_dense_engine': 0}):
        m = build(CoNet, g, dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8], **cfg))
        m.zero_grad()
        loss = m.calculate_loss(batch)
        loss.backward()
        torch.cuda.synchronize()
        res.append((m.dense_engine, loss.detach().cpu(), {n: p.grad.cpu() for n, p in m.named_parameters() if p.grad is not None}))
    assert [r[0] for r in res] == [1, 0]
    torch.testing.assert_close(res[0][1], res[1][1], rtol=1e-5, atol=0)
    differs = False
    for n in res[0][2]:
        scale = max(1e-6, float(res[1][2][n].abs().max()))
        torch.testing.assert_close(res[0][2][n], res[1][2][n], rtol=2e-4, atol=5e-5 * scale, msg=lambda s: f'grad {n}: {s}')
        differs = differs or not torch.equal(res[0][2][n], res[1][2][n])
    assert differs      # (the engines really differ: bf16x6 on tensor cores vs fp32 FMA)


@pytest.mark.parametrize('fused', [True, False])
def test_dtcdr(fused):
    from recbole_cdr_b200.model.cross_domain_recommender.dtcdr import DTCDR
    g = Golden('dtcdr_neumf')
    m = build(DTCDR, g, dict(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF',
                             alpha=g.meta('alpha'), xdr_fused_mlp=fused))
    assert m._fused_ok() == fused and m.use_fused_mlp == fused
    batch = cuda_batch(g)
    check_loss_and_grads(m, g, batch, grad_rtol=2e-4, grad_atol=2e-6)
    torch.testing.assert_close(m.predict(batch).cpu(), g.t('predict'), rtol=1e-4, atol=1e-6)


def test_seeded_construction_draws_the_reference_weights():
    """Same construction + init order as the reference => torch.manual_seed(2022) reproduces its initial weights."""
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden('emcdr_bpr_source')
    torch.manual_seed(2022)
    m = EMCDR(base_config(device='cpu', **dict(EMCDR_CFG, latent_factor_model='BPR', mapping_function='non_linear')),
              FakeDataset.from_golden(g))
    for name, p in m.named_parameters():
        assert torch.equal(p.detach(), g.param(name)), name
