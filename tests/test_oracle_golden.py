"""Pins oracle/cdr_oracle.py against tests/golden/*.npz -- outputs of the UNMODIFIED reference model classes
(EMCDR, CMF, CoNet, DTCDR, BiTGCF from /root/reference) produced by oracle/make_golden.py.  CPU only."""
import pytest
import torch

from oracle import cdr_oracle as O
from golden_util import Golden, emcdr_mapping_params, conet_params, dtcdr_params, bitgcf_graph

RTOL, ATOL = 2e-6, 1e-8  # same fp32 torch ops on the same CPU: differences are summation-order noise at most


def close(a, b, rtol=RTOL, atol=ATOL):
    torch.testing.assert_close(a.reshape(-1), b.reshape(-1), rtol=rtol, atol=atol)


def leafs(ts):
    return [t.clone().requires_grad_(True) for t in ts]


def check_emcdr_bpr(g, phase, reg_weight):
    ut, it = leafs([g.param(f'{phase}_user_embedding.weight'), g.param(f'{phase}_item_embedding.weight')])
    loss = O.emcdr_bpr_loss(ut, it, g.batch(f'{phase}_user_id'), g.batch(f'{phase}_item_id'),
                            g.batch(f'neg_{phase}_item_id'), reg_weight)
    assert loss.shape == (1,)
    close(loss, g.losses()[0])
    gu, gi = O.grads_of(loss, [ut, it])
    close(gu, g.grad(f'{phase}_user_embedding.weight'))
    close(gi, g.grad(f'{phase}_item_embedding.weight'))
    other = 'target' if phase == 'source' else 'source'
    assert not g.grad(f'{other}_user_embedding.weight').any()
    # predict (emcdr.py:179-190) is the plain dot score in SOURCE/TARGET phases
    close(O.dot_score(ut, it, g.batch(f'{phase}_user_id'), g.batch(f'{phase}_item_id')).detach(), g.t('predict'))


@pytest.mark.parametrize('phase', ['source', 'target'])
def test_emcdr_bpr(phase):
    g = Golden(f'emcdr_bpr_{phase}')
    check_emcdr_bpr(g, phase, g.meta('reg_weight'))


def check_emcdr_mf(g, phase, reg_weight):
    ut, it = leafs([g.param(f'{phase}_user_embedding.weight'), g.param(f'{phase}_item_embedding.weight')])
    loss = O.emcdr_mf_loss(ut, it, g.batch(f'{phase}_user_id'), g.batch(f'{phase}_item_id'),
                           g.batch(f'{phase}_label'), reg_weight)
    close(loss, g.losses()[0])
    gu, gi = O.grads_of(loss, [ut, it])
    close(gu, g.grad(f'{phase}_user_embedding.weight'))
    close(gi, g.grad(f'{phase}_item_embedding.weight'))


@pytest.mark.parametrize('phase', ['source', 'target'])
def test_emcdr_mf(phase):
    g = Golden(f'emcdr_mf_{phase}')
    check_emcdr_mf(g, phase, g.meta('reg_weight'))


def check_emcdr_map_and_predict(g, kind):
    ws, bs, wn, bn = emcdr_mapping_params(g)
    src, tgt = leafs([g.param(f'source_{kind}_embedding.weight'), g.param(f'target_{kind}_embedding.weight')])
    ws = leafs(ws)
    bs = [None if b is None else b.clone().requires_grad_(True) for b in bs]
    idx = g.batch('overlap')
    assert idx.dim() == 2 and idx.shape[1] == 1  # the reference's [b, 1] overlap batch (dataset.py:696)
    loss = O.emcdr_map_loss(src, tgt, idx, ws, bs)
    close(loss, g.losses()[0])
    params = [src, tgt] + ws + [b for b in bs if b is not None]
    names = [f'source_{kind}_embedding.weight', f'target_{kind}_embedding.weight'] + wn + [n for n in bn if n]
    for gr, nm in zip(O.grads_of(loss, params), names):
        close(gr, g.grad(nm))
    tabs = g.tables()
    u, i = g.t('pbatch/target_user_id'), g.t('pbatch/target_item_id')
    with torch.no_grad():
        if kind == 'user':
            pred = O.emcdr_predict_overlap_users(tabs['source_user'], tabs['target_user'], tabs['target_item'], u, i,
                                                 g.meta('n_ov_u'), ws, bs)
        else:
            pred = O.emcdr_predict_overlap_items(tabs['target_user'], tabs['source_item'], tabs['target_item'], u, i,
                                                 g.meta('n_ov_i'), ws, bs)
    close(pred, g.t('predict_overlap_phase'))
    return ws, bs


@pytest.mark.parametrize('case', ['non_linear', 'linear', 'items'])
def test_emcdr_map_and_predict(case):
    check_emcdr_map_and_predict(Golden(f'emcdr_map_{case}'), 'item' if case == 'items' else 'user')


def check_cmf(g, alpha, lam, gamma):
    ut, it = leafs([g.param('user_embedding.weight'), g.param('item_embedding.weight')])
    b = g.batch
    loss = O.cmf_loss(ut, it, b('source_user_id'), b('source_item_id'), b('source_label'), b('target_user_id'),
                      b('target_item_id'), b('target_label'), alpha, lam, gamma)
    close(loss, g.losses()[0])
    gu, gi = O.grads_of(loss, [ut, it])
    close(gu, g.grad('user_embedding.weight'))
    close(gi, g.grad('item_embedding.weight'))
    close(torch.sigmoid(O.dot_score(ut, it, b('target_user_id'), b('target_item_id'))).detach(), g.t('predict'))


def test_cmf():
    g = Golden('cmf_both')
    check_cmf(g, g.meta('alpha'), g.meta('lambda'), g.meta('gamma'))


def check_conet(g, ov_users):
    p, names = conet_params(g)
    tabs = {k: v.clone().requires_grad_(True) for k, v in g.tables().items()}
    for k in ('ws', 'bs', 'wt', 'bt', 'h'):
        p[k] = leafs(p[k])
    for k in ('out_s_w', 'out_s_b', 'out_t_w', 'out_t_b'):
        p[k] = p[k].clone().requires_grad_(True)
    b = g.batch
    n_ov = g.meta('n_ov_u') if ov_users else g.meta('n_ov_i')
    loss = O.conet_loss(tabs, p, b('source_user_id'), b('source_item_id'), b('source_label'), b('target_user_id'),
                        b('target_item_id'), b('target_label'), ov_users, n_ov)
    close(loss, g.losses()[0])
    flat, flat_names = [], []
    for k, v in tabs.items():
        flat.append(v)
        flat_names.append(f'{k}_embedding.weight')
    for k in ('ws', 'bs', 'wt', 'bt', 'h'):
        flat += p[k]
        flat_names += names[k]
    for k in ('out_s_w', 'out_s_b', 'out_t_w', 'out_t_b'):
        flat.append(p[k])
        flat_names.append(names[k])
    for gr, nm in zip(O.grads_of(loss, flat), flat_names):
        close(gr, g.grad(nm), rtol=1e-5, atol=1e-7)
    with torch.no_grad():
        close(O.conet_predict(tabs, p, b('target_user_id'), b('target_item_id')), g.t('predict'))
    return tabs, p


@pytest.mark.parametrize('tag', ['users', 'items'])
def test_conet(tag):
    check_conet(Golden(f'conet_{tag}'), tag == 'users')


def check_dtcdr(g, alpha):
    p, names = dtcdr_params(g)
    tabs = {k: v.clone().requires_grad_(True) for k, v in g.tables().items()}
    assert all(torch.isfinite(v).all() for v in tabs.values())  # the -inf fill of dtcdr.py:54-59 is overwritten by init
    for k in list(p):
        p[k] = leafs(p[k]) if isinstance(p[k], list) else p[k].clone().requires_grad_(True)
    b = g.batch
    loss = O.dtcdr_loss(tabs, p, b('source_user_id'), b('source_item_id'), b('source_label'), b('target_user_id'),
                        b('target_item_id'), b('target_label'), alpha)
    close(loss, g.losses()[0])
    flat, flat_names = [], []
    for k, v in tabs.items():
        flat.append(v)
        flat_names.append(f'{k}_embedding.weight')
    for k in p:
        if isinstance(p[k], list):
            flat += p[k]
            flat_names += names[k]
        else:
            flat.append(p[k])
            flat_names.append(names[k])
    for gr, nm in zip(O.grads_of(loss, flat), flat_names):
        close(gr, g.grad(nm), rtol=1e-5, atol=1e-7)
    with torch.no_grad():
        close(O.dtcdr_neumf_forward(tabs, b('target_user_id'), b('target_item_id'), p['t_mlp_w'], p['t_mlp_b'],
                                    p['t_out_w'], p['t_out_b']), g.t('predict'))


def test_dtcdr():
    g = Golden('dtcdr_neumf')
    check_dtcdr(g, g.meta('alpha'))


def check_bitgcf(g, way, n_layers, lam_s, lam_t, reg_weight):
    n_users, n_items, edges, deg = bitgcf_graph(g)
    adj = {d: O.bitgcf_norm_adj(edges[d][0], edges[d][1], n_users, n_items) for d in edges}
    tabs = {k: v.clone().requires_grad_(True) for k, v in g.tables().items()}
    b = g.batch
    kw = dict(n_layers=n_layers, connect_way=way, n_users=n_users, n_items=n_items, n_ov_users=g.meta('n_ov_u'),
              n_ov_items=g.meta('n_ov_i'), lam_s=lam_s, lam_t=lam_t, deg=deg, reg_weight=reg_weight)
    ls, lt = O.bitgcf_loss(tabs, adj['source'], adj['target'], b('source_user_id'), b('source_item_id'),
                           b('source_label'), b('target_user_id'), b('target_item_id'), b('target_label'), **kw)
    assert ls.shape == (1,) and lt.shape == (1,)
    close(ls, g.losses()[0])
    close(lt, g.losses()[1])
    order = list(tabs)
    for gr, k in zip(O.grads_of(ls + lt, [tabs[k] for k in order]), order):
        close(gr, g.grad(f'{k}_embedding.weight'), rtol=1e-5, atol=1e-7)
    with torch.no_grad():
        _, _, ftu, fti = O.bitgcf_forward(tabs, adj['source'], adj['target'], n_layers, way, n_users, n_items,
                                          g.meta('n_ov_u'), g.meta('n_ov_i'), lam_s, lam_t, deg)
        close((ftu[b('target_user_id')] * fti[b('target_item_id')]).sum(1), g.t('predict'))
        if g.has('full_sort_predict'):   # bitgcf.py:264-272
            fu = g.t('fbatch/target_user_id')
            n_tgt_items = g.meta('n_ov_i') + g.meta('n_tgt_i')
            close(ftu[fu] @ fti[:n_tgt_items].t(), g.t('full_sort_predict'), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('way', ['concat', 'mean'])
def test_bitgcf(way):
    check_bitgcf(Golden(f'bitgcf_{way}'), way, 2, 0.8, 0.7, 0.001)
