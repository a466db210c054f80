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
