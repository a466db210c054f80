"""tc5_mlp.cu (the EMCDR map step with all six products of a 128-row tile on tcgen05.mma kind::f16 / bf16x3, weight-gradient
accumulators resident in tensor memory) under the CPU CTA emulator, against the oracle and against the reference goldens
through the drop-in EMCDR class.  The emulator models this repository's reading of the tcgen05 descriptors (tc5.cuh), so
these tests prove the kernel's logic -- tile views, pipeline order, accumulation over tiles, epilogues -- not the reading."""
import ctypes

import numpy as np
import pytest
import torch

import emu_util
from oracle import cdr_oracle as O
from golden_util import Golden


def run(D, B, n_rows, seed, sms, backward=True, idx=None, scale=1.0, grad_loss=1.0, sched_seed=None):
    rng = np.random.RandomState(seed)
    src = (rng.randn(n_rows, D) * 0.5).astype(np.float32)
    tgt = (rng.randn(n_rows, D) * 0.5).astype(np.float32)
    W1 = (rng.randn(128, D) / np.sqrt(D)).astype(np.float32)
    b1 = (rng.randn(128) * 0.1).astype(np.float32)
    W2 = (rng.randn(D, 128) / np.sqrt(128)).astype(np.float32)
    b2 = (rng.randn(D) * 0.1).astype(np.float32)
    idx = rng.randint(0, n_rows, B).astype(np.int64) if idx is None else np.asarray(idx, np.int64)
    L = emu_util.lib()
    emu_util.config(sms=sms, seed=seed if sched_seed is None else sched_seed)
    p = emu_util.p
    dW, db = [np.zeros_like(W1), np.zeros_like(W2)], [np.zeros_like(b1), np.zeros_like(b2)]
    dsrc, dtgt = np.zeros_like(src), np.zeros_like(tgt)
    out8 = np.full(8, np.nan, np.float32)
    g = np.array([grad_loss], np.float32)
    dims = (ctypes.c_int * 3)(D, 128, D)
    ws = emu_util.workspace()
    rc = L.xdr_tc5_mlp_step(ctypes.c_int(2), dims, emu_util.ptr_array([W1, W2]), emu_util.ptr_array([b1, b2]),
                            emu_util.ptr_array(dW), emu_util.ptr_array(db), ctypes.c_int(2), ctypes.c_int(0), ctypes.c_int(0),
                            p(src), None, None, None, p(tgt), ctypes.c_int64(n_rows), ctypes.c_int64(0), ctypes.c_int(D),
                            p(idx), None, None, ctypes.c_int64(len(idx)), ctypes.c_int(1 if backward else 0), p(g),
                            ctypes.c_float(scale), p(dsrc), None, None, None, p(dtgt), None, p(out8), p(ws), None, None)
    assert rc == 0, L.emu_last_error()
    assert not ws[:64].any(), 'kernel left the workspace ticket dirty'
    # oracle
    ts = [torch.from_numpy(x).clone().requires_grad_(True) for x in (src, tgt, W1, b1, W2, b2)]
    loss = O.emcdr_map_loss(ts[0], ts[1], torch.from_numpy(idx).reshape(-1, 1), [ts[2], ts[4]], [ts[3], ts[5]]) * grad_loss
    grads = torch.autograd.grad(loss, ts)
    got = dict(loss=float(out8[0]), dsrc=dsrc, dtgt=dtgt, dW1=dW[0], db1=db[0], dW2=dW[1], db2=db[1])
    ref = dict(loss=float(loss.detach()) / grad_loss, dsrc=grads[0].numpy() * scale, dtgt=grads[1].numpy() * scale, dW1=grads[2].numpy(),
               db1=grads[3].numpy(), dW2=grads[4].numpy(), db2=grads[5].numpy())
    return got, ref


def check(got, ref, backward=True):
    assert abs(got['loss'] - ref['loss']) <= 1e-4 * abs(ref['loss'])
    if not backward:
        for k in ('dsrc', 'dtgt', 'dW1', 'dW2', 'db1', 'db2'):
            assert not got[k].any()
        return
    for k in ('dsrc', 'dtgt', 'dW1', 'dW2', 'db1', 'db2'):
        scale = np.abs(ref[k]).max()
        np.testing.assert_allclose(got[k], ref[k], rtol=2e-4, atol=2e-5 * scale, err_msg=k)


@pytest.mark.parametrize('D,B,sms', [(64, 128, 1), (64, 1, 1), (64, 127, 2), (64, 129, 1), (64, 300, 1), (64, 700, 2),
                                     (32, 200, 1), (16, 130, 3), (48, 256, 1)])
def test_map_step_matches_oracle(D, B, sms):
    """One tile, ragged tiles, several tiles per CTA (the tensor-memory weight-gradient accumulators carry over), several CTAs."""
    got, ref = run(D, B, 500, seed=B + D, sms=sms)
    check(got, ref)


def test_duplicates_scale_and_upstream_gradient():
    idx = np.r_[np.full(90, 7), np.arange(60)]            # a hot row: its gradient is a long sum of scatter-adds
    got, ref = run(64, len(idx), 100, seed=3, sms=1, idx=idx, scale=-0.5, grad_loss=3.0)
    check(got, ref)


def test_forward_only_leaves_gradients_untouched():
    got, ref = run(64, 200, 300, seed=5, sms=2, backward=False)
    check(got, ref, backward=False)


def test_result_does_not_depend_on_the_schedule():
    a, _ = run(64, 260, 300, seed=1, sms=2)
    b, _ = run(64, 260, 300, seed=1, sms=2, sched_seed=12345)
    for k in ('dW1', 'dW2', 'db1'):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-5, atol=1e-6)


def test_unsupported_stacks_are_refused():
    L = emu_util.lib()
    for dims in ([64, 64, 64], [64, 128, 32], [24, 128, 24], [128, 128, 128]):
        arr = (ctypes.c_int * 3)(*dims)
        assert L.xdr_tc5_mlp_supported(ctypes.c_int(2), arr) == 0
    assert L.xdr_tc5_mlp_supported(ctypes.c_int(2), (ctypes.c_int * 3)(64, 128, 64)) == 1
    assert L.xdr_tc5_mlp_supported(ctypes.c_int(3), (ctypes.c_int * 4)(64, 128, 64, 64)) == 0


@pytest.mark.parametrize('case', ['non_linear', 'items'])
def test_emcdr_map_phase_vs_reference_golden(case):
    """The drop-in EMCDR class with ``xdr_fused_mlp: 'tc5'`` against the unmodified reference's loss and gradients."""
    import test_emu_models as M
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden(f'emcdr_map_{case}')
    with emu_util.patched_ops():
        m = M.build_cpu(EMCDR, g, dict(M.EMCDR_CFG, latent_factor_model='BPR', mapping_function='non_linear', xdr_fused_mlp='tc5'))
        m.set_phase('OVERLAP')
        assert m.fused_mlp_engine == 'tc5'
        before = emu_util.counters()
        M.check(m, g, M.cpu_batch(g), grad_rtol=2e-4, grad_atol=2e-6)
        after = emu_util.counters(reset=False)
        assert after['umma_bf16'] > 0, 'the tcgen05 kernel did not run'
