#!/usr/bin/env python
"""Generate tests/golden/f4_*.npz by EXECUTING THE UNMODIFIED REFERENCE CLASSES of the five remaining models
(SURVEY.md section 8 F4): CLFM, DeepAPF, SSCDR, NATR, DCDCSR.  (test infrastructure; same rules as make_golden.py)

    python oracle/make_golden_f4.py

Stored per case: size scalars, the interaction edges the fake dataset was built from (SSCDR / NATR / DCDCSR read
interaction lists and history matrices from the dataset), every parameter, the batch, the reference's loss, every parameter
gradient, ``predict`` scores and -- for the models that draw with NumPy's global RNG inside ``calculate_loss`` (SSCDR's
``sample``, DCDCSR's map batch) -- the ``np.random.seed`` the harness set right before the call.
Harness-side compatibility: none of the reference arithmetic is touched; the dataset stand-in follows
``data/dataset.py:188-262`` (``get_history_matrix``) for the history matrices.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('XDR_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'recbole_shim'))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np
import torch

from recbole_cdr.model.cross_domain_recommender.clfm import CLFM  # noqa: E402
from recbole_cdr.model.cross_domain_recommender.deepapf import DeepAPF  # noqa: E402
from recbole_cdr.model.cross_domain_recommender.sscdr import SSCDR  # noqa: E402
from recbole_cdr.model.cross_domain_recommender.natr import NATR  # noqa: E402
from recbole_cdr.model.cross_domain_recommender.dcdcsr import DCDCSR  # noqa: E402
from fake_data import FakeDatasetF4, base_config, make_batch  # noqa: E402  (the stand-in the tests rebuild from the golden)
from make_golden import run_and_pack, save, sizes_scalar  # noqa: E402

D = 64


def dataset(n_ov_u, n_tgt_u, n_src_u, n_ov_i, n_tgt_i, n_src_i, seed=0, per_user=4):
    return FakeDatasetF4.random(n_ov_u, n_tgt_u, n_src_u, n_ov_i, n_tgt_i, n_src_i, seed=seed, per_user=per_user)


def with_edges(extra, ds):
    for dom in ('source', 'target'):
        extra[f'edges/{dom}_row'], extra[f'edges/{dom}_col'] = ds.edges[dom]
    return extra


def full_sort(model, ds, n_users=6):
    """``full_sort_predict`` of the model in its current phase for a few target users (its own RandomState: the batches of
    the case are drawn from the same stream as before this was added)."""
    users, _ = ds.valid_ids('target')
    fb = {'target_user_id': torch.from_numpy(np.random.RandomState(97).choice(users, n_users)).long()}
    with torch.no_grad():
        score = model.full_sort_predict(fb)
    return {'fbatch/target_user_id': fb['target_user_id'].numpy(), 'full_sort_predict': score.numpy()}


def both_batch(ds, rng, bs=96, bt=80, pairwise=False):
    b = make_batch(ds, 'source', bs, rng, pairwise=pairwise)
    b.update(make_batch(ds, 'target', bt, rng, pairwise=pairwise))
    return b


def main():
    cfg0 = lambda **kw: base_config(device='cpu', **kw)
    ds_b = dataset(25, 30, 28, 21, 30, 26)       # users and items overlap
    ds_u = dataset(41, 30, 35, 1, 50, 60)        # user overlap only
    ds_i = dataset(1, 40, 45, 31, 30, 33)        # item overlap only

    # ---------------- CLFM ----------------
    torch.manual_seed(2022)
    m = CLFM(cfg0(user_embedding_size=D, source_item_embedding_size=D, target_item_embedding_size=D, share_embedding_size=32,
                  alpha=0.3, reg_weight=1e-2), ds_b)
    batch = both_batch(ds_b, np.random.RandomState(31))
    extra = sizes_scalar(ds_b)
    extra.update({'meta/alpha': 0.3, 'meta/reg_weight': 1e-2})
    extra.update(full_sort(m, ds_b))
    save('f4_clfm', run_and_pack(m, batch, extra))

    # ---------------- DeepAPF (user overlap / item overlap) ----------------
    for tag, ds in (('users', ds_u), ('items', ds_i)):
        torch.manual_seed(2022)
        m = DeepAPF(cfg0(embedding_size=D, beta=0.5), ds)
        batch = both_batch(ds, np.random.RandomState(37))
        save(f'f4_deepapf_{tag}', run_and_pack(m, batch, sizes_scalar(ds)))

    # ---------------- SSCDR: SOURCE / TARGET triplet phases, OVERLAP map phase (users), predict in the OVERLAP phase -------
    for phase, dom in (('SOURCE', 'source'), ('TARGET', 'target')):
        torch.manual_seed(2022)
        m = SSCDR(cfg0(embedding_size=D, margin=1, mlp_hidden_size=[128], **{'lambda': 0.25}), ds_u)
        m.set_phase(phase)
        batch = make_batch(ds_u, dom, 96, np.random.RandomState(41), pairwise=True)
        extra = with_edges(sizes_scalar(ds_u), ds_u)
        if phase == 'TARGET':
            extra.update(full_sort(m, ds_u))
        save(f'f4_sscdr_{phase.lower()}', run_and_pack(m, batch, extra))
    for tag, ds, n_ov in (('users', ds_u, ds_u.num_overlap_user), ('items', ds_i, ds_i.num_overlap_item)):
        torch.manual_seed(2022)
        m = SSCDR(cfg0(embedding_size=D, margin=1, mlp_hidden_size=[128], **{'lambda': 0.25}), ds)
        m.set_phase('OVERLAP')
        rng = np.random.RandomState(43)
        idx = torch.from_numpy(rng.permutation(n_ov)[:24].reshape(-1, 1)).long()
        pb = make_batch(ds, 'target', 80, rng)
        with torch.no_grad():
            pred = m.predict(pb).numpy()
        extra = with_edges(sizes_scalar(ds), ds)
        extra.update({'pbatch/' + k: v.numpy() for k, v in pb.items()})
        extra['predict_overlap_phase'] = pred
        extra.update(full_sort(m, ds))
        extra['meta/np_seed'] = 4242
        np.random.seed(4242)
        save(f'f4_sscdr_map_{tag}', run_and_pack(m, {'overlap': idx}, extra, predict=False))

    # ---------------- NATR: phase 1 (SOURCE) and phase 2 (TARGET), item overlap and user overlap ----------------
    for tag, ds in (('items', ds_i), ('users', ds_u)):
        for phase, dom in (('SOURCE', 'source'), ('TARGET', 'target')):
            torch.manual_seed(2022)
            m = NATR(cfg0(source_embedding_size=D, target_embedding_size=D, reg_weight=1e-3, max_inter_length=3), ds)
            m.set_phase(phase)
            batch = make_batch(ds, dom, 96, np.random.RandomState(47))
            extra = with_edges(sizes_scalar(ds), ds)
            extra['meta/max_inter_length'] = 3
            save(f'f4_natr_{tag}_{phase.lower()}', run_and_pack(m, batch, extra))

    # ---------------- DCDCSR: SOURCE#1, TARGET#1, BOTH (benchmark + map loss), TARGET#2 (affine embedding) ----------------
    for tag, ds in (('users', ds_u), ('items', ds_i)):
        torch.manual_seed(2022)
        m = DCDCSR(cfg0(latent_factor_model='BPR', embedding_size=D, mlp_hidden_size=[128], k=5, map_batch_size=64), ds)
        base = with_edges(sizes_scalar(ds), ds)
        m.set_phase('SOURCE')
        save(f'f4_dcdcsr_{tag}_source1',
             run_and_pack(m, make_batch(ds, 'source', 96, np.random.RandomState(53), pairwise=True), dict(base)))
        m.set_phase('TARGET')
        extra = dict(base)
        extra.update(full_sort(m, ds))
        save(f'f4_dcdcsr_{tag}_target1',
             run_and_pack(m, make_batch(ds, 'target', 96, np.random.RandomState(59), pairwise=True), extra))
        m.set_phase('BOTH')
        extra = dict(base)
        extra['benchmark_embedding'] = m.benchmark_embedding.detach().numpy().copy()
        extra['meta/np_seed'] = 777
        np.random.seed(777)
        save(f'f4_dcdcsr_{tag}_both', run_and_pack(m, {}, extra, predict=False))
        m.set_phase('TARGET')
        extra = dict(base)
        extra['affine_embedding'] = m.affine_embedding.detach().numpy().copy()
        extra.update(full_sort(m, ds))
        save(f'f4_dcdcsr_{tag}_target2',
             run_and_pack(m, make_batch(ds, 'target', 96, np.random.RandomState(61), pairwise=True), extra))


if __name__ == '__main__':
    main()
