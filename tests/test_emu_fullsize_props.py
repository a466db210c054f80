"""The size-independent properties of tests/fullsize_props.py at toy size through the CTA emulator: checks the CHECKER
(written while no GPU was reachable) against a kernel that is parity-green on hardware.  The full-size run is
tests/test_gpu_z_fullsize.py."""
import pytest

import emu_util
import fullsize_props as P
from oracle import cdr_oracle as O


@pytest.mark.parametrize('nu,ni,dim,K,B,sms', [(500, 700, 64, 4, 96, 3), (60, 50, 32, 3, 64, 2)])
def test_train_step_properties(nu, ni, dim, K, B, sms):
    with emu_util.patched_ops(sms=sms) as ops:
        P.check_train_step_properties(ops, 'cpu', nu, ni, dim, K, B, seed=3, oracle=O)


def test_negative_sampler_property_checker_at_toy_size():
    """Same body as the full-size sampler test, small: keeps that test's own logic (used-pair lookup) exercised on CPU."""
    import numpy as np
    from oracle import sampler_oracle as S
    from recbole_cdr_b200.sampler import TargetDomainSampler
    nu, ni = 400, 300
    rng = np.random.RandomState(0)
    u, i = rng.randint(1, nu, 6000), rng.randint(1, ni, 6000)
    with emu_util.patched_ops():
        smp = TargetDomainSampler(nu, ni, u, i, device='cpu', seed=7)
        keys = rng.randint(1, nu, 2048)
        got = smp.sample_by_user_ids(keys, None, 1).numpy()
    rowptr, col = S.build_used_csr(u, i, nu)
    want, exhausted = S.neg_sample_uniform(keys, 1, rowptr, col, ni, 0, ni - 1, 7, 1)
    assert not exhausted and np.array_equal(got, want)
    used = np.unique(u.astype(np.int64) * ni + i)
    drawn = keys.astype(np.int64) * ni + got
    pos = np.searchsorted(used, drawn)
    assert not ((pos < len(used)) & (used[np.minimum(pos, len(used) - 1)] == drawn)).any()


@pytest.mark.parametrize('zipf', [None, 1.3])
def test_spmm_properties_at_toy_size(zipf):
    from recbole_cdr_b200.graph import NormAdj
    with emu_util.patched_ops(sms=3):
        P.check_spmm_properties(lambda r, c, nu, ni, device: NormAdj(r, c, nu, ni, device, chunk=16), 'cpu', 300, 200, 3000,
                                32, seed=2, zipf=zipf)
