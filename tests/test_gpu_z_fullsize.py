"""BASELINE.json configs[1] at FULL size on the GPU (EMCDR, 1M x 1M rows per domain, dim 64, B = 8192): the size-independent
properties of tests/fullsize_props.py (support of the scatter, zero column sums of the item gradients, user column sums,
linearity in the scale, additivity over launches with bit-equal per-step losses) plus direct parity of every per-step loss
and of both full gradient tables against the oracle's plain-torch arithmetic run on the same device.  (The file name sorts
last on purpose: it is the longest GPU test.)  The checker itself runs at toy size through the emulator in
tests/test_emu_fullsize_props.py."""
import pytest
import torch

import fullsize_props as P
from oracle import cdr_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('nu,ni,dim,K,B', [(1_000_000, 1_000_000, 64, 20, 8192), (300_000, 200_000, 128, 6, 8192)])
def test_train_step_properties_at_full_size(nu, ni, dim, K, B):
    from recbole_cdr_b200 import ops
    assert ops.train_steps_supported(B, dim, True)
    P.check_train_step_properties(ops, torch.device('cuda', 0), nu, ni, dim, K, B, seed=1, oracle=O)
    torch.cuda.synchronize()


def test_negative_sampler_at_full_size():
    """A18 at the headline's id-space size (1M users x 1M items, 5M interactions, 32 batches of 8192 keys in one call):
    bit-exact against the numpy oracle, every draw a valid item the user has not used."""
    import numpy as np
    from oracle import sampler_oracle as S
    from recbole_cdr_b200.sampler import TargetDomainSampler
    nu = ni = 1_000_000
    rng = np.random.RandomState(0)
    u, i = rng.randint(1, nu, 5_000_000), rng.randint(1, ni, 5_000_000)
    smp = TargetDomainSampler(nu, ni, u, i, device='cuda', seed=7)
    keys = rng.randint(1, nu, 32 * 8192)
    got = smp.sample_by_user_ids(keys, None, 1).cpu().numpy()
    rowptr, col = S.build_used_csr(u, i, nu)
    want, exhausted = S.neg_sample_uniform(keys, 1, rowptr, col, ni, 0, ni - 1, 7, 1)
    assert not exhausted
    assert np.array_equal(got, want)
    assert got.min() >= 1 and got.max() < ni
    used = np.unique(u.astype(np.int64) * ni + i)
    drawn = keys.astype(np.int64) * ni + got
    pos = np.searchsorted(used, drawn)
    hit = (pos < len(used)) & (used[np.minimum(pos, len(used) - 1)] == drawn)
    assert not hit.any()


@pytest.mark.parametrize('zipf', [None, 1.2])
def test_spmm_properties_at_bench_size(zipf):
    """A9-A10 at the size scripts/bench_bitgcf.py runs (BASELINE configs[3] at scale 0.25: 750k users x 500k items, ~8.8M
    interactions per domain, dim 64); the Zipf case concentrates edges on a few items (split-row path)."""
    from recbole_cdr_b200.graph import NormAdj
    P.check_spmm_properties(NormAdj, torch.device('cuda', 0), 750_000, 500_000, 8_800_000, 64, seed=4, zipf=zipf)
    torch.cuda.synchronize()
