"""End-to-end ``trainer.fit`` of every drop-in model on CPU through the CTA emulator: the host batch multiplexer
(``CrossDomainDataloader`` in its four states), the device negative sampler, the phase loop of ``CrossDomainTrainer`` /
``DCDCSRTrainer`` and the models' ``set_phase`` sequences, with the kernels' real sources underneath.  The reference's own
test style (tests/test_model.py:10-11: an epoch per phase must run) plus what it never asserts: finite, falling losses."""
import numpy as np
import pytest
import torch

import emu_util
from fake_data import FakeDatasetF4, base_config


def make_world(pairwise, overlap='users', seed=0, batch=64):
    from recbole_cdr_b200.data import CrossDomainDataloader, DomainTrainDataLoader, OverlapDataloader
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler, TargetDomainSampler
    if overlap == 'users':
        ds = FakeDatasetF4.random(31, 40, 36, 1, 50, 45, seed=seed, per_user=5)
        n_ov = ds.num_overlap_user
    elif overlap == 'items':
        ds = FakeDatasetF4.random(1, 40, 36, 27, 40, 35, seed=seed, per_user=5)
        n_ov = ds.num_overlap_item
    else:
        ds = FakeDatasetF4.random(21, 30, 28, 17, 30, 26, seed=seed, per_user=5)
        n_ov = ds.num_overlap_user
    s_u, s_i = ds.edges['source']
    t_u, t_i = ds.edges['target']
    s_smp = CrossDomainSourceSampler('train', ds, user_ids=s_u, item_ids=s_i, device='cpu').set_phase('train')
    t_smp = TargetDomainSampler(ds.num_total_user, ds.target_domain_dataset.num('target_item_id'), t_u, t_i, device='cpu')
    g = torch.Generator().manual_seed(seed)
    src = DomainTrainDataLoader('source_user_id', 'source_item_id', s_u, s_i, batch, s_smp, pairwise, 'source_label',
                                shuffle=True, generator=g)
    tgt = DomainTrainDataLoader('target_user_id', 'target_item_id', t_u, t_i, batch, t_smp, pairwise, 'target_label',
                                shuffle=True, generator=g)
    return ds, CrossDomainDataloader(src, tgt, OverlapDataloader(n_ov, 16, True, g))


CASES = [
    # name, pairwise, overlap layout, model config, train_modes, epochs, phases whose loss must fall
    ('EMCDR', True, 'users', dict(latent_factor_model='BPR', source_embedding_size=64, target_embedding_size=64, reg_weight=0.01,
                                  mapping_function='non_linear', mlp_hidden_size=[128]), ['SOURCE', 'TARGET', 'OVERLAP'], 2, None),
    ('CMF', False, 'both', dict(embedding_size=64, alpha=0.5, gamma=0.0, **{'lambda': 0.0}), ['BOTH'], 2, None),
    ('CoNet', False, 'users', dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8], xdr_dense_engine=0), ['BOTH'], 2, None),
    ('CoNet', False, 'users', dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8], xdr_fused_conet=True),
     ['BOTH'], 2, None),
    ('DTCDR', False, 'both', dict(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF', alpha=0.5,
                                  xdr_fused_mlp='tc'), ['BOTH'], 2, None),
    ('CLFM', False, 'both', dict(user_embedding_size=64, source_item_embedding_size=64, target_item_embedding_size=64,
                                 share_embedding_size=32, alpha=0.5, reg_weight=1e-4), ['BOTH'], 2, None),
    ('DeepAPF', False, 'items', dict(embedding_size=64, beta=0.5), ['BOTH'], 2, None),
    ('SSCDR', True, 'users', {'embedding_size': 64, 'margin': 1, 'mlp_hidden_size': [128], 'lambda': 0.25},
     ['SOURCE', 'TARGET', 'OVERLAP'], 2, ('OVERLAP',)),
    ('NATR', False, 'items', dict(source_embedding_size=64, target_embedding_size=64, reg_weight=1e-3, max_inter_length=4),
     ['SOURCE', 'TARGET'], 2, None),
    ('DCDCSR', True, 'users', dict(latent_factor_model='BPR', embedding_size=64, mlp_hidden_size=[128], k=5, map_batch_size=32),
     ['SOURCE', 'TARGET', 'BOTH', 'TARGET'], 2, ('SOURCE', 'BOTH')),
]


@pytest.mark.parametrize('name,pairwise,overlap,cfg,modes,epochs,falling', CASES,
                         ids=[f'{c[0]}{"-fused" if any(str(k).startswith("xdr_") for k in c[3]) else ""}' for c in CASES])
def test_fit_runs_every_phase(name, pairwise, overlap, cfg, modes, epochs, falling):
    from recbole_cdr_b200.utils import ModelType, get_model, get_trainer
    np.random.seed(0)
    with emu_util.patched_ops(sms=2):
        ds, loader = make_world(pairwise, overlap)
        full = base_config(device='cpu', learner='adam', learning_rate=0.01, weight_decay=0.0, train_modes=modes,
                           epoch_num=[str(epochs)] * len(modes), source_split=False, **cfg)
        torch.manual_seed(2022)
        model = get_model(name)(full, ds)
        trainer = get_trainer(ModelType.CROSSDOMAIN, name)(full, model)
        seen = []
        trainer.fit(loader, callback_fn=lambda epoch, loss: seen.append((len(seen) // epochs, epoch, loss)))
    assert len(seen) == epochs * len(modes)
    assert np.isfinite([l for _, _, l in seen]).all(), seen
    check = range(len(modes)) if falling is None else [k for k, m in enumerate(modes) if m in falling]
    for k in check:
        l = [x for p, _, x in seen if p == k]
        assert l[-1] < l[0], (name, modes[k], l)
    if name in ('EMCDR', 'SSCDR', 'DCDCSR'):
        assert model.phase == 'OVERLAP'       # trainer.py:75 / :129


def test_fit_validates_stops_early_and_checkpoints(tmp_path):
    """recbole's per-phase Trainer.fit contract around the reference's phase loop (trainer.py:59-73): every eval_step epochs the
    evaluation function is called, the best score is kept and checkpointed (reference checkpoint keys), and the phase stops after
    stopping_step validations without improvement; valid_data without an evaluation function warns instead of vanishing."""
    from recbole_cdr_b200.utils import ModelType, get_model, get_trainer
    np.random.seed(0)
    with emu_util.patched_ops(sms=2):
        ds, loader = make_world(False, 'both')
        full = base_config(device='cpu', learner='adam', learning_rate=0.01, weight_decay=0.0, train_modes=['BOTH'],
                           epoch_num=['8'], source_split=False, embedding_size=64, alpha=0.5, gamma=0.0, eval_step=1,
                           stopping_step=2, checkpoint_dir=str(tmp_path), **{'lambda': 0.0})
        torch.manual_seed(2022)
        model = get_model('CMF')(full, ds)
        trainer = get_trainer(ModelType.CROSSDOMAIN, 'CMF')(full, model)
        with pytest.warns(UserWarning, match='no evaluation function'):
            trainer.fit(loader, valid_data='held-out', saved=False)
        scores = iter([0.1, 0.3, 0.2, 0.25, 0.28, 0.9, 0.9, 0.9])     # best at the 2nd validation, then three without improvement
        calls = []

        def valid_fn(m, vd):
            assert vd == 'held-out' and not m.training
            calls.append(next(scores))
            return calls[-1], {'recall@10': calls[-1]}
        trainer.set_valid_fn(valid_fn)
        epochs = []
        best, result = trainer.fit(loader, valid_data='held-out', saved=True, callback_fn=lambda e, l: epochs.append(e))
    assert calls == [0.1, 0.3, 0.2, 0.25, 0.28] and epochs == [0, 1, 2, 3, 4]          # stopped after 3 > stopping_step misses
    assert best == 0.3 and result == {'recall@10': 0.3}
    ckpt = torch.load(trainer.saved_model_file, weights_only=False)
    assert set(ckpt) >= {'config', 'epoch', 'cur_step', 'best_valid_score', 'state_dict', 'other_parameter', 'optimizer'}
    assert ckpt['epoch'] == 1 and ckpt['best_valid_score'] == 0.3
    assert set(ckpt['state_dict']) == set(model.state_dict())
