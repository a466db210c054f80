#!/usr/bin/env python
"""Generate tests/golden/v_*.npz: a SECOND set of reference outputs for the five hot-path models (EMCDR, CMF, CoNet,
DTCDR, BiTGCF), on other shapes and hyper-parameters than make_golden.py's, again by EXECUTING THE UNMODIFIED REFERENCE
CLASSES.  (test infrastructure; same rules as make_golden.py)

    python oracle/make_golden_variants.py

What the first set does not pin and this one does: row widths other than 64 (16, 32, 128), batches that are not a multiple
of 4 or of the kernels' row tiles, heavily duplicated ids, reg_weight == 0, deeper mapping / tower stacks (incl. CoNet.yaml's
[64, 32, 16, 8]), other loss weights, BiTGCF with one and three propagation layers -- and ``full_sort_predict`` of every
model that has one (the first set stores ``predict`` only), which is what the fused score + top-k path (SURVEY.md section 8
F2) has to agree with.

Every file is self-describing: ``meta/spec_json`` holds the model name, the config keyword arguments, the phase, and the
six id-space sizes, so the tests that consume these files are table-driven (tests/variants_util.py).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import numpy as np
import torch

import make_golden as MG  # noqa: E402  (sets up sys.path for the reference, applies the NumPy-2 shim)
from make_golden import FakeDataset, base_config, make_batch, run_and_pack, save, sizes_scalar  # noqa: E402
from oracle import cdr_oracle  # noqa: E402

MODELS = {'EMCDR': MG.EMCDR, 'CMF': MG.CMF, 'CoNet': MG.CoNet, 'DTCDR': MG.DTCDR, 'BiTGCF': MG.BiTGCF}


def skewed_batch(ds, domain, B, rng, pairwise):
    """Heavily duplicated ids: two thirds of the rows come from three hot users / items."""
    b = make_batch(ds, domain, B, rng, pairwise=pairwise)
    users, items = ds.valid_ids(domain)
    hot = rng.rand(B) < 0.66
    b[f'{domain}_user_id'][hot] = torch.from_numpy(rng.choice(users[:3], int(hot.sum()))).long()
    b[f'{domain}_item_id'][hot] = torch.from_numpy(rng.choice(items[:3], int(hot.sum()))).long()
    return b


def full_sort(model, ds, rng, n_users=7):
    users, _ = ds.valid_ids('target')
    fb = {'target_user_id': torch.from_numpy(rng.choice(users, n_users)).long()}
    with torch.no_grad():
        score = model.full_sort_predict(fb)
    return {'fbatch/target_user_id': fb['target_user_id'].numpy(), 'full_sort_predict': score.numpy()}


def emit(name, model_name, cfg, sizes, phase, batch_fn, *, with_edges=False, predict=True, full_sort_pred=False,
         overlap_predict=False, ds_seed=0):
    ds = FakeDataset(*sizes, seed=ds_seed)
    torch.manual_seed(2023)
    model = MODELS[model_name](base_config(**cfg), ds)
    if phase is not None:
        model.set_phase(phase)
    rng = np.random.RandomState(sum(map(ord, name)))   # a stable per-case seed
    batch = batch_fn(ds, rng)
    extra = sizes_scalar(ds)
    extra['meta/spec_json'] = json.dumps({'model': model_name, 'cfg': cfg, 'phase': phase, 'sizes': list(sizes)})
    if with_edges:
        for dom in ('source', 'target'):
            extra[f'edges/{dom}_row'], extra[f'edges/{dom}_col'] = ds.edges[dom]
    if overlap_predict:  # EMCDR's OVERLAP-phase predict mixes mapped and plain rows: its own batch
        pb = make_batch(ds, 'target', 61, rng)
        with torch.no_grad():
            extra['predict_overlap_phase'] = model.predict(pb).numpy()
        extra.update({'pbatch/' + k: v.numpy() for k, v in pb.items()})
    if full_sort_pred:
        extra.update(full_sort(model, ds, rng))
    save(name, run_and_pack(model, batch, extra, predict=predict))


def both(bs, bt, pairwise=False, skew=False):
    def fn(ds, rng):
        mk = skewed_batch if skew else make_batch
        b = mk(ds, 'source', bs, rng, pairwise=pairwise)
        b.update(mk(ds, 'target', bt, rng, pairwise=pairwise))
        return b
    return fn


def one(domain, B, pairwise, skew=False):
    def fn(ds, rng):
        return (skewed_batch if skew else make_batch)(ds, domain, B, rng, pairwise)
    return fn


def overlap(n, kind):
    def fn(ds, rng):
        n_ov = ds.num_overlap_user if kind == 'user' else ds.num_overlap_item
        return {'overlap': torch.from_numpy(rng.permutation(n_ov)[:n].reshape(-1, 1)).long()}
    return fn


def main():
    users_only = (37, 23, 29, 1, 44, 52)    # user overlap (items disjoint)
    items_only = (1, 33, 38, 27, 25, 31)    # item overlap
    both_ov = (19, 26, 22, 17, 28, 24)      # users and items overlap

    # ---------------- EMCDR ----------------
    def emcdr_cfg(d, lfm, reg, mf='non_linear', hidden=(128,)):
        return dict(latent_factor_model=lfm, source_embedding_size=d, target_embedding_size=d, reg_weight=reg,
                    mapping_function=mf, mlp_hidden_size=list(hidden))

    emit('v_emcdr_bpr_source_d32', 'EMCDR', emcdr_cfg(32, 'BPR', 0.0), users_only, 'SOURCE', one('source', 97, True),
         full_sort_pred=False)
    emit('v_emcdr_bpr_target_d128', 'EMCDR', emcdr_cfg(128, 'BPR', 0.05), users_only, 'TARGET',
         one('target', 256, True, skew=True), full_sort_pred=True)
    emit('v_emcdr_mf_target_d16', 'EMCDR', emcdr_cfg(16, 'MF', 0.001), items_only, 'TARGET', one('target', 63, False),
         full_sort_pred=True)
    emit('v_emcdr_map_users_deep', 'EMCDR', emcdr_cfg(32, 'BPR', 0.01, hidden=(64, 48)), users_only, 'OVERLAP',
         overlap(33, 'user'), predict=False, overlap_predict=True, full_sort_pred=True)
    emit('v_emcdr_map_items_linear', 'EMCDR', emcdr_cfg(64, 'BPR', 0.01, mf='linear'), items_only, 'OVERLAP',
         overlap(27, 'item'), predict=False, overlap_predict=True, full_sort_pred=True)

    # ---------------- CMF ----------------
    emit('v_cmf_d32', 'CMF', {'embedding_size': 32, 'alpha': 0.7, 'lambda': 0.0, 'gamma': 0.1}, both_ov, None,
         both(50, 130), full_sort_pred=True)
    emit('v_cmf_d128_skewed', 'CMF', {'embedding_size': 128, 'alpha': 0.5, 'lambda': 0.2, 'gamma': 0.2}, both_ov, None,
         both(200, 64, skew=True), full_sort_pred=False)

    # ---------------- CoNet ----------------
    emit('v_conet_yaml_stack', 'CoNet', dict(embedding_size=32, reg_weight=0.1, mlp_hidden_size=[64, 32, 16, 8]), users_only,
         None, both(150, 150), full_sort_pred=True)
    emit('v_conet_items_wide', 'CoNet', dict(embedding_size=64, reg_weight=0.0, mlp_hidden_size=[64, 32]), items_only, None,
         both(77, 77, skew=True))

    # ---------------- DTCDR ----------------
    emit('v_dtcdr_deep', 'DTCDR', dict(embedding_size=32, mlp_hidden_size=[64, 32, 16], dropout_prob=0.0, base_model='NeuMF',
                                       alpha=0.1), both_ov, None, both(70, 90))
    emit('v_dtcdr_users_only', 'DTCDR', dict(embedding_size=64, mlp_hidden_size=[32], dropout_prob=0.0, base_model='NeuMF',
                                             alpha=0.9), users_only, None, both(128, 33, skew=True))

    # ---------------- BiTGCF ----------------
    MG.BiTGCF.get_norm_adj_mat = lambda self, inter, n_users=None, n_items=None: cdr_oracle.bitgcf_norm_adj(
        inter.row, inter.col, n_users, n_items)   # see make_golden.py: SciPy-compatible build, cross-checked there
    emit('v_bitgcf_l3_concat', 'BiTGCF', dict(embedding_size=64, n_layers=3, reg_weight=0.01, lambda_source=0.5,
                                              lambda_target=0.9, drop_rate=0.0, connect_way='concat'), both_ov, None,
         both(110, 75), with_edges=True, full_sort_pred=True)
    emit('v_bitgcf_l1_mean', 'BiTGCF', dict(embedding_size=16, n_layers=1, reg_weight=0.0, lambda_source=1.0,
                                            lambda_target=0.3, drop_rate=0.0, connect_way='mean'), users_only, None,
         both(64, 64, skew=True), with_edges=True, full_sort_pred=True, ds_seed=3)


if __name__ == '__main__':
    main()
