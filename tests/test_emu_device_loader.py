"""data.DeviceDomainTrainDataLoader (interactions, shuffle and negative draw resident on the device) against the host loader
it mirrors, on CPU with the sampler kernel under the emulator: the same sampler seed gives the SAME batches field by field
in all four multiplexer states, and a pointwise model (CoNet) trains an epoch from it through the unchanged trainer loop."""
import numpy as np
import pytest
import torch

import emu_util
from fake_data import FakeDataset, base_config


def loaders(ds, device_side, pairwise, batch_size=64, seed=11):
    from recbole_cdr_b200.data import DeviceDomainTrainDataLoader
    from recbole_cdr_b200.data.dataloader import CrossDomainDataloader, DomainTrainDataLoader, OverlapDataloader
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler, TargetDomainSampler
    rng = np.random.RandomState(0)
    doms = {}
    for dom in ('source', 'target'):
        users, items = ds.valid_ids(dom)
        doms[dom] = (rng.choice(users, 500), rng.choice(items, 500))
    s_smp = CrossDomainSourceSampler('train', ds, user_ids=doms['source'][0], item_ids=doms['source'][1], device='cpu',
                                     seed=seed).set_phase('train')
    t_smp = TargetDomainSampler(ds.num_total_user, ds.num_overlap_item + ds.num_target_only_item, doms['target'][0],
                                doms['target'][1], device='cpu', seed=seed + 1)
    cls = DeviceDomainTrainDataLoader if device_side else DomainTrainDataLoader
    kw = dict(device='cpu') if device_side else {}
    src = cls('source_user_id', 'source_item_id', *doms['source'], batch_size, s_smp, pairwise, label_field='source_label', **kw)
    tgt = cls('target_user_id', 'target_item_id', *doms['target'][0:2], batch_size + 6, t_smp, pairwise,
              label_field='target_label', **kw)
    return CrossDomainDataloader(src, tgt, OverlapDataloader(ds.num_overlap_user, 16, generator=torch.Generator().manual_seed(1)))


@pytest.mark.parametrize('pairwise', [True, False])
def test_device_loader_yields_the_host_loaders_batches(pairwise):
    from recbole_cdr_b200.utils.enum_type import CrossDomainDataLoaderState as S
    ds = FakeDataset(41, 30, 35, 21, 50, 60)
    with emu_util.patched_ops():
        host, devl = loaders(ds, False, pairwise), loaders(ds, True, pairwise)
        for state in (S.SOURCE, S.TARGET, S.BOTH, S.OVERLAP):
            host.set_mode(state)
            devl.set_mode(state)
            a, b = list(host), list(devl)
            assert len(a) == len(b) == len(host) and len(a) > 1
            for x, y in zip(a, b):
                assert set(x.interaction) == set(y.interaction)
                for k in x.interaction:
                    assert torch.equal(x[k], y[k]), (state, k)
            lens = {v.shape[0] for v in a[-1].interaction.values()}
            assert state == S.OVERLAP or min(lens) < (64 if pairwise else 64), 'the ragged last batch is kept'


def test_conet_trains_an_epoch_from_the_device_loader():
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    from recbole_cdr_b200.trainer import CrossDomainTrainer
    from recbole_cdr_b200.utils.enum_type import CrossDomainDataLoaderState as S
    ds = FakeDataset(41, 30, 35, 1, 50, 60)
    with emu_util.patched_ops():
        torch.manual_seed(0)
        cfg = base_config(device='cpu', embedding_size=16, reg_weight=0.0, mlp_hidden_size=[16, 8], learner='adam',
                          learning_rate=0.01, weight_decay=0.0, train_modes=['BOTH'], epoch_num=['4'], source_split=False,
                          xdr_dense_engine=0)   # (loader / trainer logic: the fp32 tiles emulate 5x faster than tcgen05)
        model = CoNet(cfg, ds)
        trainer = CrossDomainTrainer(cfg, model)
        data = loaders(ds, True, False)
        data.set_mode(S.BOTH)
        losses = [trainer._train_epoch(data, 0) for _ in range(4)]
        assert all(np.isfinite(losses)) and losses[-1] < losses[0]
