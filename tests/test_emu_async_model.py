"""The CTA emulator's asynchronous models have teeth.  A tiny tcgen05 kernel (tests/emu/emu_kernels.cpp ``async_probe_kernel``,
test infrastructure only) computes D = A B^T correctly, and with one protocol mistake injected at a time -- each of which is
undefined behaviour or a race on the hardware -- it must NOT: tensor memory read without waiting for the commit, operand stores
that no fence.proxy.async followed, tcgen05.ld results used before tcgen05.wait::ld, an operand plane overwritten before the MMA
has executed.  So "emulator-green" for the tcgen05 kernels means their waits and fences are in place, not only their math."""
import ctypes

import numpy as np
import pytest

import emu_util


def probe(fault, seed=0):
    L = emu_util.lib()
    emu_util.config(sms=1, seed=seed)
    rng = np.random.RandomState(3)
    A = rng.randn(128, 16).astype(np.float32)
    B = rng.randn(16, 16).astype(np.float32)
    D = np.zeros((128, 16), np.float32)
    p = emu_util.p
    assert L.emu_async_probe(p(A), p(B), p(D), ctypes.c_int(fault)) == 0
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    mass = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    with np.errstate(invalid='ignore'):
        return bool(np.all(np.abs(D - ref) <= 2.0 ** -14 * mass + 1e-6))


@pytest.mark.parametrize('seed', [0, 7])
def test_correct_protocol_gives_the_product(seed):
    assert probe(0, seed)


@pytest.mark.parametrize('fault,what', [(1, 'TMEM read without waiting for the commit'), (2, 'operand stores without a proxy fence'),
                                        (3, 'tcgen05.ld results used before wait::ld'),
                                        (4, 'operand plane overwritten before the MMA executed')])
def test_protocol_mistakes_are_caught(fault, what):
    assert not probe(fault), what
