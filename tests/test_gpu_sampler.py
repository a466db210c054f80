"""GPU parity of xdr_neg_sample_uniform with oracle/sampler_oracle.py: integer work, so BIT-EXACT."""
import numpy as np
import pytest
import torch

from oracle import sampler_oracle as S
from fake_data import FakeDataset

pytestmark = pytest.mark.gpu


def source_case(seed, n_inter=4000):
    ds = FakeDataset(61, 50, 70, 21, 40, 60)
    rng = np.random.RandomState(seed)
    users, items = ds.valid_ids('source')
    u = rng.choice(users, n_inter)
    i = rng.choice(items, n_inter)
    return ds, u, i


@pytest.mark.parametrize('num', [1, 3])
@pytest.mark.parametrize('n_keys', [1, 257, 5000])
def test_source_sampler_bit_exact_and_contract(num, n_keys):
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler
    ds, u, i = source_case(0)
    smp = CrossDomainSourceSampler('train', ds, user_ids=u, item_ids=i, device='cuda', seed=99).set_phase('train')
    keys = np.random.RandomState(1).choice(u, n_keys)
    got = smp.sample_by_user_ids(torch.from_numpy(keys), None, num)
    assert got.is_cuda and got.dtype == torch.int64 and got.shape == (n_keys * num,)
    rowptr, col = S.build_used_csr(u, i, ds.num_total_user)
    want, exhausted = S.neg_sample_uniform(keys, num, rowptr, col, ds.num_overlap_item, ds.num_target_only_item,
                                           (ds.num_overlap_item - 1) + ds.num_source_only_item, 99, smp._calls)
    assert not exhausted
    assert np.array_equal(got.cpu().numpy(), want)                       # bit-exact
    assert torch.equal(smp.used_rowptr.cpu(), torch.from_numpy(rowptr)) and torch.equal(smp.used_col.cpu(), torch.from_numpy(col))
    valid = set(ds.valid_ids('source')[1].tolist())
    used = {(a, b) for a, b in zip(u.tolist(), i.tolist())}
    for k, v in zip(np.tile(keys, num).tolist(), got.cpu().tolist()):
        assert v in valid and (k, v) not in used


def test_target_sampler_and_errors():
    from recbole_cdr_b200.sampler import TargetDomainSampler
    rng = np.random.RandomState(3)
    u, i = rng.randint(1, 200, 3000), rng.randint(1, 300, 3000)
    smp = TargetDomainSampler(200, 300, u, i, device='cuda', seed=5)
    keys = rng.randint(1, 200, 1000)
    got = smp.sample_by_user_ids(keys, None, 2).cpu().numpy()
    rowptr, col = S.build_used_csr(u, i, 200)
    want, _ = S.neg_sample_uniform(keys, 2, rowptr, col, 300, 0, 299, 5, 1)
    assert np.array_equal(got, want) and got.min() >= 1 and got.max() < 300
    with pytest.raises(ValueError, match='not exist'):
        smp.sample_by_user_ids(np.array([5, 200]), None, 1)               # key outside [0, n_users)
    with pytest.raises(ValueError, match='all items'):
        TargetDomainSampler(3, 4, [1, 1, 1], [1, 2, 3], device='cuda')      # a user that used every item
    with pytest.raises(NotImplementedError):
        TargetDomainSampler(3, 4, [1], [1], device='cuda', distribution='popularity')
