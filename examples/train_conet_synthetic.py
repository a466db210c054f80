#!/usr/bin/env python
"""CoNet on synthetic data on one B200: BOTH-mode steps (conet.py:183-203) through the drop-in class -- one stacked pass over
the source and the target batch, cross-stitch layers on the tcgen05 dense engine -- replayed as a CUDA graph with SGD inside.

    python examples/train_conet_synthetic.py [--scale 100000] [--batch 16384] [--steps 200] [--engine 1] [--eager]

--engine  dense engine of the cross-stitch layers: 1 tcgen05 (default), 0 fp32 FMA tiles, 2 tcgen05 where it wins per call
--eager   launch every kernel from Python instead of replaying the captured step (what the trainer's plain inner loop does)
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'recbole-cdr_b200'))

import torch

from recbole_cdr_b200.data import Interaction, synthetic
from recbole_cdr_b200.data.idspace import IdSpace
from recbole_cdr_b200.utils import get_model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scale', type=int, default=100_000, help='users per domain (BASELINE config #3: 5_000_000 users, 2_000_000 items)')
    ap.add_argument('--batch', type=int, default=16384, help='rows per domain and step (positives then negatives, labels 1 / 0)')
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--dim', type=int, default=128)
    ap.add_argument('--engine', type=int, default=1, choices=[0, 1, 2])
    ap.add_argument('--lr', type=float, default=0.05)
    ap.add_argument('--eager', action='store_true')
    a = ap.parse_args()
    emu = os.environ.get('XDR_EXAMPLE_EMU') == '1'   # developer aid: run the kernels' sources under the CPU emulator of tests/emu
    if emu:
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        sys.path.insert(0, ROOT)
        import emu_util
        ctx = emu_util.patched_ops(sms=2)     # (kept referenced: the patch lasts as long as the context object)
        ctx.__enter__()
        a.eager = True      # (no CUDA graphs on the CPU)
    else:
        assert torch.cuda.is_available(), 'this example needs a CUDA device (the hot path has no CPU implementation)'
    dev = 'cpu' if emu else 'cuda'
    half, n_items = a.scale // 2, max(2, 2 * a.scale // 5)
    # user-overlap scenario (half of each domain's users overlapped), items disjoint: joint id layout of data/dataset.py:344-445
    ds = synthetic.SyntheticCrossDomainDataset(IdSpace(half + 1, half, half), IdSpace(1, n_items, n_items))
    cfg = {'source_domain': {'NEG_PREFIX': 'neg_'}, 'target_domain': {'NEG_PREFIX': 'neg_'}, 'device': dev,
           'embedding_size': a.dim, 'reg_weight': 0.01, 'mlp_hidden_size': [64, 32, 16, 8], 'xdr_dense_engine': a.engine}
    torch.manual_seed(2022)
    model = get_model('CoNet')(cfg, ds).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=a.lr)

    def batch(step):
        b = synthetic.make_batch(ds, 'source', a.batch, 1 + 2 * step, dev, pairwise=False)
        b.update(synthetic.make_batch(ds, 'target', a.batch, 2 + 2 * step, dev, pairwise=False))
        return Interaction(b)

    # a learnable signal: label = whether user and item ids have the same parity (random labels would only fit noise)
    def relabel(b):
        for d in ('source', 'target'):
            b[f'{d}_label'] = ((b[f'{d}_user_id'] + b[f'{d}_item_id']) % 2 == 0).float()
        return b

    if a.eager:
        def step(b):
            opt.zero_grad(set_to_none=True)
            loss = model.calculate_loss(b)
            loss.backward()
            opt.step()
            return loss.detach()
    else:
        from recbole_cdr_b200.trainer import GraphedTrainStep
        graphed = GraphedTrainStep(model, relabel(batch(0)), optimizer=opt)   # forward + backward + SGD in one CUDA graph

        def step(b):
            graphed.zero_table_grads()
            return graphed(b)
    reg = float(sum(p.weight.detach().norm() for p in model.crossparas))
    print(f'CoNet {ds.num_total_user} users x {ds.num_total_item} items, dim {a.dim}, 2 x {a.batch} rows per step, '
          f'engine {model.dense_engine}, {"eager" if a.eager else "graph replay"}; sum_l ||H_l||_F at init = {reg:.3f}')
    t0, shown = time.time(), max(1, a.steps // 10)
    for s in range(a.steps):
        show = s % shown == 0 or s == a.steps - 1
        if show:    # the regulariser's share of the loss this step is about to report (weights before the update)
            reg = float(sum(p.weight.detach().norm() for p in model.crossparas))
        loss = step(relabel(batch(s)))
        if show:
            if not emu:
                torch.cuda.synchronize()
            print(f'step {s:5d}: loss {float(loss):.4f}  (BCE part {float(loss) - reg:.4f})   {time.time() - t0:.2f} s')
    probs = model.predict(relabel(batch(a.steps)))
    print('target-tower predictions of a fresh batch: mean %.3f, min %.3f, max %.3f' % (float(probs.mean()), float(probs.min()),
                                                                                         float(probs.max())))


if __name__ == '__main__':
    main()
