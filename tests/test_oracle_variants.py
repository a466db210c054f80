"""Pins oracle/cdr_oracle.py against the SECOND set of reference outputs (tests/golden/v_*.npz, produced by the unmodified
reference classes through oracle/make_golden_variants.py): other row widths, ragged batches, duplicated ids, deeper stacks,
reg_weight 0, 1- and 3-layer BiTGCF -- and ``full_sort_predict``, restated here from the reference line by line.  CPU only."""
import pytest
import torch

import variants_util as V
from golden_util import Golden
from oracle import cdr_oracle as O
from test_oracle_golden import (check_bitgcf, check_cmf, check_conet, check_dtcdr, check_emcdr_bpr, check_emcdr_map_and_predict,
                                check_emcdr_mf, close)


def n_target_items(g):
    return g.meta('n_ov_i') + g.meta('n_tgt_i')


def emcdr(g, sp):
    cfg, phase = sp['cfg'], sp['phase']
    tabs = g.tables()
    n_t = n_target_items(g)
    fu = g.t('fbatch/target_user_id') if g.has('full_sort_predict') else None
    if phase in ('SOURCE', 'TARGET'):
        dom = phase.lower()
        (check_emcdr_bpr if cfg['latent_factor_model'] == 'BPR' else check_emcdr_mf)(g, dom, cfg['reg_weight'])
        if fu is not None:   # emcdr.py:216-219
            assert phase == 'TARGET'
            close(tabs['target_user'][fu] @ tabs['target_item'][:n_t].t(), g.t('full_sort_predict'))
        return
    kind = 'item' if g.meta('n_ov_u') == 1 else 'user'
    ws, bs = check_emcdr_map_and_predict(g, kind)
    ws, bs = [w.detach() for w in ws], [None if b is None else b.detach() for b in bs]
    if fu is None:
        return
    if kind == 'user':       # emcdr.py:221-226: mapped source rows for overlapped users, plain target rows otherwise
        mapped = O.emcdr_mapping(tabs['source_user'][fu], ws, bs)
        user_e = torch.where((fu < g.meta('n_ov_u')).unsqueeze(1), mapped, tabs['target_user'][fu])
        items = tabs['target_item'][:n_t]
    else:                    # emcdr.py:227-231: the overlapped item block is replaced by its mapped source rows
        user_e = tabs['target_user'][fu]
        n_ov = g.meta('n_ov_i')
        items = torch.cat([O.emcdr_mapping(tabs['source_item'][:n_ov], ws, bs), tabs['target_item'][n_ov:n_t]], dim=0)
    close(user_e @ items.t(), g.t('full_sort_predict'), rtol=1e-5, atol=1e-7)


def cmf(g, sp):
    cfg = sp['cfg']
    check_cmf(g, cfg['alpha'], cfg['lambda'], cfg['gamma'])
    if g.has('full_sort_predict'):   # cmf.py:107-112: raw dot products, no sigmoid
        fu = g.t('fbatch/target_user_id')
        close(g.param('user_embedding.weight')[fu] @ g.param('item_embedding.weight')[:n_target_items(g)].t(),
              g.t('full_sort_predict'))


def conet(g, sp):
    tabs, p = check_conet(g, g.meta('n_ov_i') == 1)
    if g.has('full_sort_predict'):   # conet.py:222-244: the target tower of predict() on every (user, target item) pair
        fu = g.t('fbatch/target_user_id')
        n_t = n_target_items(g)
        users = fu.repeat_interleave(n_t)
        items = torch.arange(n_t).repeat(fu.numel())
        with torch.no_grad():
            close(O.conet_predict(tabs, p, users, items), g.t('full_sort_predict'), rtol=1e-5, atol=1e-7)


def dtcdr(g, sp):
    check_dtcdr(g, sp['cfg']['alpha'])


def bitgcf(g, sp):
    cfg = sp['cfg']
    check_bitgcf(g, cfg['connect_way'], cfg['n_layers'], cfg['lambda_source'], cfg['lambda_target'], cfg['reg_weight'])


DISPATCH = {'EMCDR': emcdr, 'CMF': cmf, 'CoNet': conet, 'DTCDR': dtcdr, 'BiTGCF': bitgcf}


def test_the_variant_table_is_complete():
    assert len(V.VARIANTS) == V.EXPECTED, V.VARIANTS
    assert {V.spec(Golden(n))['model'] for n in V.VARIANTS} == set(DISPATCH)


@pytest.mark.parametrize('name', V.VARIANTS)
def test_oracle_matches_the_reference_on_variant(name):
    g = Golden(name)
    sp = V.spec(g)
    DISPATCH[sp['model']](g, sp)
