#!/usr/bin/env python
"""EMCDR on synthetic data, end to end on one B200: the three training phases of the reference (SOURCE, TARGET, OVERLAP)
through the drop-in classes, then the fused full-sort top-k.

    python examples/train_emcdr_synthetic.py [--scale 100000] [--epochs 2] [--row-optimizer adagrad] [--tc]

--row-optimizer  step the embedding tables with the row-sparse optimizer kernel over the batch's ids only
--tc             run the OVERLAP (mapping) phase through the tensor-core fused MLP kernel (xdr_fused_mlp: 'tc')
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'recbole-cdr_b200'))

import numpy as np
import torch

from recbole_cdr_b200.data import CrossDomainDataloader, DomainTrainDataLoader, Interaction, OverlapDataloader, synthetic
from recbole_cdr_b200.sampler import CrossDomainSourceSampler, TargetDomainSampler
from recbole_cdr_b200.utils import ModelType, get_model, get_trainer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scale', type=int, default=100_000, help='users / items per domain (BASELINE config #2 uses 1_000_000)')
    ap.add_argument('--interactions', type=int, default=2_000_000)
    ap.add_argument('--batch', type=int, default=8192)
    ap.add_argument('--epochs', type=int, default=2)
    ap.add_argument('--row-optimizer', choices=['sgd', 'adagrad', 'lazy_adam'], default=None)
    ap.add_argument('--tc', action='store_true')
    a = ap.parse_args()
    emu = os.environ.get('XDR_EXAMPLE_EMU') == '1'   # developer aid: run the kernels' sources under the CPU emulator of tests/emu
    if emu:
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        sys.path.insert(0, ROOT)
        import emu_util
        ctx = emu_util.patched_ops(sms=2)
        ctx.__enter__()
    else:
        assert torch.cuda.is_available(), 'this example needs a CUDA device (the hot path has no CPU implementation)'
    dev = 'cpu' if emu else 'cuda'
    ds = synthetic.emcdr_scale(a.scale)                        # joint id layout of data/dataset.py:344-445
    rng = np.random.RandomState(0)
    inter = {}
    for dom in ('source', 'target'):
        b = synthetic.make_batch(ds, dom, a.interactions, 1 if dom == 'source' else 2, 'cpu', pairwise=False)
        inter[dom] = (b[f'{dom}_user_id'].numpy(), b[f'{dom}_item_id'].numpy())
    s_smp = CrossDomainSourceSampler('train', ds, user_ids=inter['source'][0], item_ids=inter['source'][1], device=dev).set_phase('train')
    t_smp = TargetDomainSampler(ds.num_total_user, ds.target_domain_dataset.num('target_item_id'), *inter['target'], device=dev)
    g = torch.Generator().manual_seed(0)
    loaders = [DomainTrainDataLoader(f'{d}_user_id', f'{d}_item_id', *inter[d], a.batch, smp, True, f'{d}_label', shuffle=True,
                                     generator=g) for d, smp in (('source', s_smp), ('target', t_smp))]
    loader = CrossDomainDataloader(loaders[0], loaders[1], OverlapDataloader(ds.num_overlap_user, a.batch, True, g))
    cfg = {'source_domain': {'NEG_PREFIX': 'neg_'}, 'target_domain': {'NEG_PREFIX': 'neg_'}, 'device': dev,
           'latent_factor_model': 'BPR', 'source_embedding_size': 64, 'target_embedding_size': 64, 'reg_weight': 0.01,
           'mapping_function': 'non_linear', 'mlp_hidden_size': [128], 'learner': 'adam', 'learning_rate': 0.01,
           'weight_decay': 0.0, 'train_modes': ['SOURCE', 'TARGET', 'OVERLAP'], 'epoch_num': [str(a.epochs)] * 3,
           'source_split': False}
    if a.row_optimizer:
        cfg['xdr_row_optimizer'] = a.row_optimizer
    if a.tc:
        cfg['xdr_fused_mlp'] = 'tc'
    torch.manual_seed(2022)
    model = get_model('EMCDR')(cfg, ds).to(dev)
    trainer = get_trainer(ModelType.CROSSDOMAIN, 'EMCDR')(cfg, model)
    t0 = [time.time()]

    def report(epoch, loss):
        if not emu:
            torch.cuda.synchronize()
        print(f'phase {model.phase:8s} epoch {epoch}: summed loss {loss:.4f}   ({time.time() - t0[0]:.2f} s)')
        t0[0] = time.time()

    trainer.fit(loader, callback_fn=report)
    users = torch.arange(1, 9, device=dev)
    scores, items = model.full_sort_topk(Interaction({'target_user_id': users}), 10)
    print('top-10 target items of users 1..8 (OVERLAP phase: mapped user vectors):')
    print(items.cpu().numpy())


if __name__ == '__main__':
    main()
