"""Table-driven access to tests/golden/v_*.npz (oracle/make_golden_variants.py): outputs of the unmodified reference classes
on a second set of shapes / hyper-parameters, each file carrying its own spec (model, config kwargs, phase)."""
import glob
import importlib
import json
import os

import torch

from fake_data import FakeDataset, base_config
from golden_util import GOLDEN_DIR

VARIANTS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'v_*.npz')))
EXPECTED = 13   # keep in step with make_golden_variants.py: a missing fixture must fail, not shrink the table


def spec(g):
    return json.loads(str(g.z['meta/spec_json']))


def model_class(name):
    return getattr(importlib.import_module(f'recbole_cdr_b200.model.cross_domain_recommender.{name.lower()}'), name)


def build(g, device, **cfg_extra):
    """The drop-in class of the golden's model on `device`, loaded with the golden's parameters, in the golden's phase."""
    sp = spec(g)
    edges = None
    if g.has('edges/source_row'):
        edges = {dom: (g.z[f'edges/{dom}_row'], g.z[f'edges/{dom}_col']) for dom in ('source', 'target')}
    ds = FakeDataset.from_golden(g, edges)
    torch.manual_seed(0)
    cfg = dict(sp['cfg'])
    cfg.update(cfg_extra)
    m = model_class(sp['model'])(base_config(device=device, **cfg), ds)
    m.load_state_dict({n: g.param(n) for n in g.param_names()}, strict=True)
    m = m.to(device)
    if sp['phase']:
        m.set_phase(sp['phase'])
    return m


def batch(g, device, prefix='batch/'):
    from recbole_cdr_b200.data import Interaction
    return Interaction({k[len(prefix):]: torch.from_numpy(g.z[k]).to(device) for k in g.z.files if k.startswith(prefix)})


def check_against_reference(m, g, device, loss_rtol=1e-4, grad_rtol=2e-4, grad_atol=2e-6, pred_rtol=1e-4, pred_atol=2e-6):
    """loss, every parameter gradient, predict, OVERLAP-phase predict and full_sort_predict against the reference's."""
    b = batch(g, device)
    m.zero_grad()
    loss = m.calculate_loss(b)
    losses = list(loss) if isinstance(loss, tuple) else [loss]
    assert len(losses) == len(g.losses())
    for got, ref in zip(losses, g.losses()):
        torch.testing.assert_close(got.detach().reshape(-1).cpu(), ref.reshape(-1), rtol=loss_rtol, atol=0)
    sum(l.sum() for l in losses).backward()
    for name, p in m.named_parameters():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        ref = g.grad(name)
        # rows hit by many duplicate ids are long fp32 sums in another order: scale the floor with the table's largest entry
        atol = max(grad_atol, 2e-6 * float(ref.abs().max()))
        torch.testing.assert_close(got.cpu(), ref, rtol=grad_rtol, atol=atol, msg=lambda s: f'grad {name}: {s}')
    with torch.no_grad():
        if g.has('predict'):
            torch.testing.assert_close(m.predict(b).cpu(), g.t('predict'), rtol=pred_rtol, atol=pred_atol)
        if g.has('predict_overlap_phase'):
            torch.testing.assert_close(m.predict(batch(g, device, 'pbatch/')).cpu(), g.t('predict_overlap_phase'),
                                       rtol=pred_rtol, atol=pred_atol)
        if g.has('full_sort_predict'):
            got = m.full_sort_predict(batch(g, device, 'fbatch/'))
            torch.testing.assert_close(got.cpu().reshape(-1), g.t('full_sort_predict').reshape(-1), rtol=pred_rtol,
                                       atol=pred_atol)


def check_topk_against_reference(m, g, device, k=5):
    """``full_sort_topk`` (fused scoring + masking + top-k) against torch.topk of the REFERENCE's full_sort_predict scores
    with the PAD column and a per-user history masked the way recbole's full-sort evaluation does."""
    import numpy as np
    fb = batch(g, device, 'fbatch/')
    n_users = fb['target_user_id'].numel()
    ref = g.t('full_sort_predict').reshape(n_users, -1).clone()
    rng = np.random.RandomState(5)
    ptr, ids = [0], []
    for _ in range(n_users):
        h = np.unique(rng.randint(1, ref.shape[1], 6))
        ids.append(h)
        ptr.append(ptr[-1] + len(h))
    hp, hi = torch.tensor(ptr), torch.from_numpy(np.concatenate(ids))
    ref[:, 0] = -float('inf')
    for u in range(n_users):
        ref[u, hi[hp[u]:hp[u + 1]]] = -float('inf')
    rs, ri = torch.topk(ref, k, dim=1)
    sc, pos = m.full_sort_topk(fb, k, hp.to(device), hi.to(device))
    torch.testing.assert_close(sc.cpu(), rs, rtol=1e-4, atol=2e-6)
    # positions may swap only where the reference's own scores tie to within the tolerance
    same = pos.cpu() == ri
    if not bool(same.all()):
        alt = torch.gather(ref, 1, pos.cpu())
        torch.testing.assert_close(alt, rs, rtol=1e-4, atol=2e-6)
