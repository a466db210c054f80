"""The device negative sampler (neg_sample.cu, bit-exact on a B200 against oracle/sampler_oracle.py) through the CPU CTA
emulator: integer work, so bit-exact here too.  Pins the emulator's integer / atomic paths (Philox, 64-bit mulhi, atomicOr)."""
import numpy as np
import pytest
import torch

import emu_util
from fake_data import FakeDataset
from oracle import sampler_oracle as S


@pytest.mark.parametrize('num,n_keys', [(1, 1), (3, 257), (2, 1500)])
def test_source_sampler_bit_exact(num, n_keys):
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler
    ds = FakeDataset(61, 50, 70, 21, 40, 60)
    rng = np.random.RandomState(0)
    users, items = ds.valid_ids('source')
    u, i = rng.choice(users, 4000), rng.choice(items, 4000)
    with emu_util.patched_ops():
        smp = CrossDomainSourceSampler('train', ds, user_ids=u, item_ids=i, device='cpu', seed=99).set_phase('train')
        keys = np.random.RandomState(1).choice(u, n_keys)
        got = smp.sample_by_user_ids(torch.from_numpy(keys), None, num)
    rowptr, col = S.build_used_csr(u, i, ds.num_total_user)
    want, exhausted = S.neg_sample_uniform(keys, num, rowptr, col, ds.num_overlap_item, ds.num_target_only_item,
                                           (ds.num_overlap_item - 1) + ds.num_source_only_item, 99, smp._calls)
    assert not exhausted
    assert np.array_equal(got.numpy(), want)


def test_unknown_key_is_reported():
    from recbole_cdr_b200.sampler import TargetDomainSampler
    rng = np.random.RandomState(3)
    u, i = rng.randint(1, 200, 3000), rng.randint(1, 300, 3000)
    with emu_util.patched_ops():
        smp = TargetDomainSampler(200, 300, u, i, device='cpu', seed=5)
        with pytest.raises(ValueError, match='not exist'):
            smp.sample_by_user_ids(np.array([5, 200]), None, 1)
