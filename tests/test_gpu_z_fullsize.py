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
