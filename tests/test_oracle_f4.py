"""Pins oracle/f4_oracle.py against tests/golden/f4_*.npz -- outputs of the UNMODIFIED reference classes CLFM, DeepAPF,
SSCDR, NATR, DCDCSR (oracle/make_golden_f4.py).  CPU only."""
import numpy as np
import pytest
import torch

from fake_data import FakeDatasetF4
from golden_util import Golden
from oracle import f4_oracle as F4

RTOL, ATOL = 5e-6, 1e-8


def params(g):
    return {n: g.param(n).clone().requires_grad_(True) for n in g.param_names()}


def check(loss, P, g, rtol=RTOL):
    torch.testing.assert_close(loss.detach().reshape(-1), g.losses()[0].reshape(-1), rtol=rtol, atol=ATOL)
    loss.sum().backward()
    for n, p in P.items():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        torch.testing.assert_close(got, g.grad(n), rtol=1e-5, atol=1e-8, msg=lambda s: f'{n}: {s}')


def test_clfm():
    g = Golden('f4_clfm')
    P = params(g)
    b = lambda k: g.batch(k)
    loss = F4.clfm_loss(P, b('source_user_id'), b('source_item_id'), b('source_label'), b('target_user_id'),
                        b('target_item_id'), b('target_label'), g.meta('alpha'), g.meta('reg_weight'))
    check(loss, P, g)
    torch.testing.assert_close(F4.clfm_forward(P, 'target', b('target_user_id'), b('target_item_id')).detach(), g.t('predict'),
                               rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_deepapf(tag):
    g = Golden(f'f4_deepapf_{tag}')
    P = params(g)
    b = lambda k: g.batch(k)
    ov_users = tag == 'users'
    n_ov = g.meta('n_ov_u') if ov_users else g.meta('n_ov_i')
    check(F4.deepapf_loss(P, b('source_user_id'), b('source_item_id'), b('source_label'), b('target_user_id'),
                          b('target_item_id'), b('target_label'), ov_users, n_ov), P, g)


@pytest.mark.parametrize('domain', ['source', 'target'])
def test_sscdr_rec(domain):
    g = Golden(f'f4_sscdr_{domain}')
    P = params(g)
    check(F4.sscdr_rec_loss(P, domain, g.batch(f'{domain}_user_id'), g.batch(f'{domain}_item_id'),
                            g.batch(f'neg_{domain}_item_id'), 1), P, g)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_sscdr_map(tag):
    """The draws come from the drop-in class's host sampler under the golden's seed (its RNG call order is the
    reference's); the oracle then restates the arithmetic."""
    from fake_data import base_config
    from recbole_cdr_b200.model.cross_domain_recommender.sscdr import SSCDR
    g = Golden(f'f4_sscdr_map_{tag}')
    P = params(g)
    m = SSCDR(base_config(device='cpu', embedding_size=64, margin=1, mlp_hidden_size=[128], **{'lambda': 0.25}),
              FakeDatasetF4.from_golden(g))
    idx = g.batch('overlap').squeeze(1)
    np.random.seed(g.meta('np_seed'))
    pos, neg = m.sample(idx, mode='user' if tag == 'users' else 'item')
    check(F4.sscdr_map_loss(P, idx, pos, neg, tag == 'users', 1, 0.25), P, g)


@pytest.mark.parametrize('tag', ['items', 'users'])
def test_natr(tag):
    g = Golden(f'f4_natr_{tag}_source')
    P = params(g)
    check(F4.natr_phase1_loss(P, g.batch('source_user_id'), g.batch('source_item_id'), g.batch('source_label')), P, g)
    g = Golden(f'f4_natr_{tag}_target')
    P = params(g)
    for k in ('source_user_embedding.weight', 'source_item_embedding.weight'):
        P[k] = P[k].detach()                                      # frozen in phase 2 (natr.py:79-83)
    ds = FakeDatasetF4.from_golden(g)
    hist, _, lens = ds.history_item_matrix(domain='target') if tag == 'items' else ds.history_user_matrix(domain='target')
    hist = hist[:, :g.meta('max_inter_length')]
    mask = (torch.arange(hist.shape[1]) < lens.unsqueeze(1)).float()
    loss = F4.natr_phase2_loss(P, g.batch('target_user_id'), g.batch('target_item_id'), g.batch('target_label'),
                               tag == 'items', hist, mask, 1e-3)
    check(loss, {k: v for k, v in P.items() if v.requires_grad}, g)


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_dcdcsr(tag):
    side = 'user' if tag == 'users' else 'item'
    for stage, dom in (('source1', 'source'), ('target1', 'target')):
        g = Golden(f'f4_dcdcsr_{tag}_{stage}')
        P = params(g)
        check(F4.dcdcsr_rec_loss(P[f'{dom}_user_embedding.weight'], P[f'{dom}_item_embedding.weight'], g.batch(f'{dom}_user_id'),
                                 g.batch(f'{dom}_item_id'), g.batch(f'neg_{dom}_item_id')), P, g)
    g = Golden(f'f4_dcdcsr_{tag}_both')
    P = params(g)
    ds = FakeDatasetF4.from_golden(g)
    hist = ds.history_item_matrix if tag == 'users' else ds.history_user_matrix
    pop_s, pop_t = hist(domain='source')[2].float(), hist(domain='target')[2].float()
    n_ov = g.meta('n_ov_u') if tag == 'users' else g.meta('n_ov_i')
    n_tgt = n_ov + (g.meta('n_tgt_u') if tag == 'users' else g.meta('n_tgt_i'))
    with torch.no_grad():
        bench = F4.dcdcsr_benchmark(P[f'source_{side}_embedding.weight'][:n_ov], P[f'target_{side}_embedding.weight'], pop_s,
                                    pop_t, 5)
    torch.testing.assert_close(bench, g.t('benchmark_embedding'), rtol=1e-5, atol=1e-7)
    np.random.seed(g.meta('np_seed'))
    sampled = torch.from_numpy(np.random.randint(0, n_tgt, 64))
    check(F4.dcdcsr_map_loss(P, P[f'target_{side}_embedding.weight'], bench, sampled), P, g)
    g = Golden(f'f4_dcdcsr_{tag}_target2')
    P = params(g)
    affine = g.t('affine_embedding')
    ut, it = (affine, P['target_item_embedding.weight']) if tag == 'users' else (P['target_user_embedding.weight'], affine)
    check(F4.dcdcsr_rec_loss(ut, it, g.batch('target_user_id'), g.batch('target_item_id'), g.batch('neg_target_item_id')), P, g)
