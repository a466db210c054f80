"""Pins oracle/optim_oracle.py against torch.optim itself (the optimizers the reference instantiates through recbole's
``_build_optimizer``), then runs the row-sparse optimizer kernel under the CTA emulator against that oracle."""
import numpy as np
import pytest
import torch

import emu_util
from oracle import optim_oracle as OO


def case(seed=0, n=60, d=16, b=40):
    rng = np.random.RandomState(seed)
    w = rng.randn(n, d).astype(np.float32) * 0.1
    ids = rng.randint(0, n, b)
    ids[:5] = ids[0]  # duplicates
    rows = rng.randn(b, d).astype(np.float32) * 0.05
    g = np.zeros_like(w)
    np.add.at(g, ids, rows)
    return w, ids, rows, g


def test_adagrad_oracle_equals_torch_dense_adagrad():
    w, ids, rows, g = case()
    p = torch.nn.Parameter(torch.from_numpy(w.copy()))
    opt = torch.optim.Adagrad([p], lr=0.05)
    s = np.zeros_like(w)
    ww = w.copy()
    for step in range(3):
        _, ids, rows, g = case(seed=step)
        p.grad = torch.from_numpy(g.copy())
        opt.step()
        ww, s = OO.adagrad_step(ww, s, g, 0.05)
        np.testing.assert_allclose(ww, p.detach().numpy(), rtol=1e-6, atol=1e-8)


def test_sparse_adam_oracle_equals_torch_sparse_adam():
    w, _, _, _ = case()
    emb = torch.nn.Embedding(w.shape[0], w.shape[1], sparse=True)
    with torch.no_grad():
        emb.weight.copy_(torch.from_numpy(w))
    opt = torch.optim.SparseAdam(list(emb.parameters()), lr=0.01)
    ww, m, v = w.copy(), np.zeros_like(w), np.zeros_like(w)
    for step in range(1, 4):
        _, ids, rows, g = case(seed=step)
        opt.zero_grad()
        out = emb(torch.from_numpy(ids))
        (out * torch.from_numpy(rows)).sum().backward()
        opt.step()
        ww, m, v = OO.sparse_adam_step(ww, m, v, g, ids, step, 0.01)
        np.testing.assert_allclose(ww, emb.weight.detach().numpy(), rtol=2e-5, atol=5e-7)  # fp32 eps at |w| ~ 0.1


def test_sgd_oracle_equals_torch_sgd():
    w, ids, rows, g = case()
    p = torch.nn.Parameter(torch.from_numpy(w.copy()))
    p.grad = torch.from_numpy(g.copy())
    torch.optim.SGD([p], lr=0.1).step()
    np.testing.assert_allclose(OO.sgd_step(w, g, 0.1), p.detach().numpy(), rtol=1e-6, atol=1e-8)


def emu_optim(kind, w, g, ids, stamp, step_id, lr, s1=None, s2=None, adam_t=1, eps=1e-8, seed=0):
    import ctypes
    L = emu_util.lib()
    emu_util.config(sms=2, seed=seed)
    oob = np.zeros(1, dtype=np.int32)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    p = emu_util.p
    rc = L.xdr_sparse_optim_rows(ctypes.c_int(kind), p(w), p(g), p(s1), p(s2), p(stamp), p(ids), ctypes.c_int64(ids.size),
                                     ctypes.c_int64(w.shape[0]), ctypes.c_int(w.shape[1]), ctypes.c_int(step_id),
                                     ctypes.c_int64(adam_t), ctypes.c_float(lr), ctypes.c_float(eps), ctypes.c_double(0.9),
                                     ctypes.c_double(0.999), p(oob), None)
    assert rc == 0, L.emu_last_error()
    emu_util.config(4, 0)
    return int(oob[0])


@pytest.mark.parametrize('seed', [0, 5])
def test_emulated_kernel_matches_oracle_over_three_steps(seed):
    """SGD, Adagrad and lazy Adam: weights and state after three steps with duplicate ids; gradient table left at zero."""
    for kind, name in ((0, 'sgd'), (1, 'adagrad'), (2, 'adam')):
        w0, _, _, _ = case()
        w = w0.copy()
        stamp = np.zeros(w.shape[0], dtype=np.int32)
        s1, s2 = np.zeros_like(w), np.zeros_like(w)
        rw, rs, rm, rv = w0.copy(), np.zeros_like(w), np.zeros_like(w), np.zeros_like(w)
        for step in range(1, 4):
            _, ids, rows, g = case(seed=10 + step)
            gk = g.copy()
            emu_optim(kind, w, gk, ids, stamp, step, 0.05, s1 if kind else None, s2 if kind == 2 else None, adam_t=step,
                      eps=1e-10 if kind == 1 else 1e-8, seed=seed)
            assert not gk.any(), 'gradient rows must be zero after the step'
            if kind == 0:
                rw = OO.sgd_step(rw, g, 0.05)
            elif kind == 1:
                rw, rs = OO.adagrad_step(rw, rs, g, 0.05)
                np.testing.assert_allclose(s1, rs, rtol=1e-6, atol=1e-12, err_msg=name)
            else:
                rw, rm, rv = OO.sparse_adam_step(rw, rm, rv, g, ids, step, 0.05)
                np.testing.assert_allclose(s1, rm, rtol=1e-5, atol=1e-10, err_msg=name)
                np.testing.assert_allclose(s2, rv, rtol=1e-5, atol=1e-12, err_msg=name)
            np.testing.assert_allclose(w, rw, rtol=1e-5, atol=1e-7, err_msg=name)  # 1e-7 ~ one fp32 ulp at |w| ~ 0.3


def test_emulated_kernel_flags_out_of_range_ids_and_skips_them():
    w0, ids, rows, g = case()
    w, gk = w0.copy(), g.copy()
    ids = ids.copy()
    ids[3] = w.shape[0] + 7
    stamp = np.zeros(w.shape[0], dtype=np.int32)
    assert emu_optim(0, w, gk, ids, stamp, 1, 0.1) == 1


def test_row_sparse_optimizer_class_under_emulator():
    """trainer.RowSparseOptimizer: two id lists on one table (positive and negative items) are one update per row."""
    from recbole_cdr_b200.trainer.row_optim import RowSparseOptimizer
    w0, ids, rows, g = case()
    with emu_util.patched_ops():
        table = torch.nn.Parameter(torch.from_numpy(w0.copy()))
        opt = RowSparseOptimizer('adagrad', lr=0.05)
        rw, rs = w0.copy(), np.zeros_like(w0)
        for step in range(1, 3):
            _, ids, rows, g = case(seed=20 + step)
            table.grad = torch.from_numpy(g.copy())
            half = len(ids) // 2
            opt.step([(table, torch.from_numpy(ids[:half])), (table, torch.from_numpy(ids[half:]))])
            rw, rs = OO.adagrad_step(rw, rs, g, 0.05)
            assert not table.grad.any()
            np.testing.assert_allclose(table.detach().numpy(), rw, rtol=1e-5, atol=1e-7)
