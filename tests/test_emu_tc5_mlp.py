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


def test_correction_pass_completes_an_eager_step():
    """backward = 2 adds (g - 1) x the gradients on top of a backward = 1 launch that used an upstream gradient of 1 -- together
    they are the step with upstream gradient g -- and is a no-op for g = 1."""
    D, B, n_rows, sms = 64, 200, 300, 2
    rng = np.random.RandomState(11)
    src, tgt = (rng.randn(n_rows, D) * 0.5).astype(np.float32), (rng.randn(n_rows, D) * 0.5).astype(np.float32)
    W1, b1 = (rng.randn(128, D) / 8).astype(np.float32), (rng.randn(128) * 0.1).astype(np.float32)
    W2, b2 = (rng.randn(D, 128) / 11).astype(np.float32), (rng.randn(D) * 0.1).astype(np.float32)
    idx = rng.randint(0, n_rows, B).astype(np.int64)
    L, p = emu_util.lib(), emu_util.p
    emu_util.config(sms=sms, seed=1)
    dims = (ctypes.c_int * 3)(D, 128, D)

    def launch(mode, g, dst):
        dW, db, dsrc, dtgt = dst
        out8, gg, ws = np.full(8, np.nan, np.float32), np.array([g], np.float32), emu_util.workspace()
        rc = L.xdr_tc5_mlp_step(ctypes.c_int(2), dims, emu_util.ptr_array([W1, W2]), emu_util.ptr_array([b1, b2]),
                                emu_util.ptr_array(dW), emu_util.ptr_array(db), ctypes.c_int(2), ctypes.c_int(0), ctypes.c_int(0),
                                p(src), None, None, None, p(tgt), ctypes.c_int64(n_rows), ctypes.c_int64(0), ctypes.c_int(D),
                                p(idx), None, None, ctypes.c_int64(B), ctypes.c_int(mode), p(gg) if g is not None else None,
                                ctypes.c_float(1.0), p(dsrc), None, None, None, p(dtgt), None, p(out8), p(ws), None, None)
        assert rc == 0, L.emu_last_error()
        return out8

    fresh = lambda: ([np.zeros_like(W1), np.zeros_like(W2)], [np.zeros_like(b1), np.zeros_like(b2)], np.zeros_like(src), np.zeros_like(tgt))
    for g in (1.0, 0.25, 3.0):
        want = fresh()
        launch(1, g, want)                       # the ordinary backward with upstream gradient g
        got = fresh()
        launch(1, None, got)                     # eager: upstream gradient 1 at forward time ...
        snap = [a.copy() for a in got[0]]
        out8 = launch(2, g, got)                 # ... then the correction
        if g == 1.0:
            assert np.isnan(out8).all(), 'the correction pass must return before it does anything when g == 1'
            for a, b in zip(snap, got[0]):
                np.testing.assert_array_equal(a, b)
        for a, b in zip(want[0] + want[1] + [want[2], want[3]], got[0] + got[1] + [got[2], got[3]]):
            np.testing.assert_allclose(b, a, rtol=1e-4, atol=1e-6 * max(1.0, float(np.abs(a).max())))
    assert L.xdr_tc5_mlp_step is not None


def test_eager_step_through_the_drop_in_class():
    """'inplace' table gradients + the tcgen05 engine: calculate_loss accumulates every gradient in ONE launch, backward() only
    corrects; same loss and gradients as the reference golden, also when the loss is scaled before backward()."""
    import test_emu_models as M
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden('emcdr_map_non_linear')
    for factor in (1.0, 0.5):
        with emu_util.patched_ops() as ops:
            prev = ops.get_table_grad_mode()
            ops.set_table_grad_mode('inplace')
            try:
                m = M.build_cpu(EMCDR, g, dict(M.EMCDR_CFG, latent_factor_model='BPR', mapping_function='non_linear'))
                m.set_phase('OVERLAP')
                assert m.fused_mlp_auto and m.fused_mlp_engine == 'tc5'
                calls = []
                real = ops.call
                ops.call = lambda nm, *a, **k: (calls.append((nm, a[21])) if nm == 'xdr_tc5_mlp_step' else None, real(nm, *a, **k))[1]
                try:
                    loss = m.calculate_loss(M.cpu_batch(g))
                    (loss * factor).sum().backward()
                finally:
                    ops.call = real
                assert [c[1] for c in calls] == [1, 2], calls      # one eager launch, one correction launch
                torch.testing.assert_close(loss.detach().reshape(-1), g.losses()[0].reshape(-1), rtol=1e-4, atol=0)
                for name, prm in m.named_parameters():
                    got = prm.grad if prm.grad is not None else torch.zeros_like(prm)
                    torch.testing.assert_close(got, g.grad(name) * factor, rtol=2e-4, atol=2e-6, msg=lambda s: f'{name} (x{factor}): {s}')
            finally:
                ops.set_table_grad_mode(prev)
