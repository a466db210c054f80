"""CPU runs of the fused gather -> MLP -> loss -> backward -> scatter kernels under the CTA emulator (tests/emu).

``impl 0`` is ``fused_mlp_kernel`` -- parity-green on a B200 (tests/test_gpu_kernels.py) -- and is run here to pin the
emulator itself; ``impl 1`` is ``tc_mlp_kernel`` (3xTF32 mma.sync tiles), written without GPU access.  Both are compared
with the oracle (``oracle/cdr_oracle.py``) at the tolerances of the GPU parity tests.  This checks kernel *logic* only;
the ``gpu`` tests remain the parity gate on hardware."""
import numpy as np
import pytest
import torch

import emu_util
from oracle import cdr_oracle as O

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
LOSS_RTOL = 1e-4


def rand_table(n, d, seed, std=0.1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g) * std


def rand_ids(n, hi, seed, zipf=None):
    rng = np.random.RandomState(seed)
    if zipf:
        return torch.from_numpy(np.minimum(rng.zipf(zipf, n) - 1, hi - 1)).long()
    return torch.from_numpy(rng.randint(0, hi, n)).long()


def close(got, want, nm, rtol=2e-4):
    want = want.detach().numpy() if isinstance(want, torch.Tensor) else want
    atol = max(1e-7, 1e-4 * float(np.abs(want).max()))
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, err_msg=nm)


def map_case(batch, hidden=128, dim=64, linear=False):
    g = torch.Generator().manual_seed(111)
    src, tgt = rand_table(300, dim, 112, 0.3), rand_table(300, dim, 113, 0.3)
    if linear:
        ws, bs = [torch.randn(dim, dim, generator=g) * 0.2], [None]
    else:
        ws = [torch.randn(hidden, dim, generator=g) * 0.2, torch.randn(dim, hidden, generator=g) * 0.2]
        bs = [torch.randn(hidden, generator=g) * 0.1, torch.randn(dim, generator=g) * 0.1]
    idx = rand_ids(batch, 300, 114, 1.3)
    leaves = [t.clone().requires_grad_(True) for t in [src, tgt] + ws + [b for b in bs if b is not None]]
    nw = len(ws)
    lb = leaves[2 + nw:] if not linear else [None]
    ref = O.emcdr_map_loss(leaves[0], leaves[1], idx.view(-1, 1), leaves[2:2 + nw], lb)
    (ref * 1.3).backward()
    return src, tgt, ws, bs, idx, leaves, ref


@pytest.mark.parametrize('impl,batch,tile_rows', [
    (0, 33, 0), (0, 100, 0),                                      # hardware-validated kernel: pins the emulator
    (1, 1, 0), (1, 31, 32), (1, 33, 32), (1, 64, 64), (1, 100, 64), (1, 200, 32), (1, 333, 0),
])
def test_map_step_matches_oracle(impl, batch, tile_rows):
    src, tgt, ws, bs, idx, leaves, ref = map_case(batch)
    emu_util.config(sms=3, seed=0)
    r = emu_util.mlp_step(impl, [64, 128, 64], [w.numpy() for w in ws], [b.numpy() for b in bs], ACT_TANH, 0, 0,
                          (src.numpy(), None, None, None, tgt.numpy()), idx.numpy(), None, None, grad_loss=1.3,
                          tile_rows=tile_rows)
    assert abs(r['loss'] - ref.item()) <= LOSS_RTOL * abs(ref.item())
    close(r['dtabs'][0], leaves[0].grad, 'src')
    close(r['dtabs'][4], leaves[1].grad, 'tgt')
    close(r['dW'][0], leaves[2].grad, 'W1')
    close(r['dW'][1], leaves[3].grad, 'W2')
    close(r['db'][0], leaves[4].grad, 'b1')
    close(r['db'][1], leaves[5].grad, 'b2')


def test_tc_map_step_linear_mapping():
    """mapping_function 'linear': one bias-free Linear(64, 64) (emcdr.py:58-59)."""
    src, tgt, ws, bs, idx, leaves, ref = map_case(77, linear=True)
    emu_util.config(sms=2, seed=0)
    r = emu_util.mlp_step(1, [64, 64], [ws[0].numpy()], [None], ACT_TANH, 0, 0,
                          (src.numpy(), None, None, None, tgt.numpy()), idx.numpy(), None, None, grad_loss=1.3)
    assert abs(r['loss'] - ref.item()) <= LOSS_RTOL * abs(ref.item())
    close(r['dtabs'][0], leaves[0].grad, 'src')
    close(r['dtabs'][4], leaves[1].grad, 'tgt')
    close(r['dW'][0], leaves[2].grad, 'W')


def test_tc_map_step_is_schedule_independent():
    """Randomised fiber scheduling: a missing __syncthreads would make the result depend on the seed."""
    src, tgt, ws, bs, idx, leaves, ref = map_case(70)
    outs = []
    for seed in (0, 1, 7):
        emu_util.config(sms=2, seed=seed)
        r = emu_util.mlp_step(1, [64, 128, 64], [w.numpy() for w in ws], [b.numpy() for b in bs], ACT_TANH, 0, 0,
                              (src.numpy(), None, None, None, tgt.numpy()), idx.numpy(), None, None, tile_rows=32)
        outs.append(r)
    emu_util.config(sms=4, seed=0)
    for r in outs[1:]:
        assert r['loss'] == outs[0]['loss']
        np.testing.assert_array_equal(r['dW'][0], outs[0]['dW'][0])
        # scatter order changes with the schedule; duplicates are summed in another order
        np.testing.assert_allclose(r['dtabs'][0], outs[0]['dtabs'][0], rtol=1e-5, atol=1e-9)


def dtcdr_case(batch, dim=64, hidden=(32, 16)):
    g = torch.Generator().manual_seed(211)
    tabs = {k: rand_table(150, dim, 212 + i, 0.3) for i, k in enumerate(('source_user', 'target_user', 'source_item',
                                                                         'target_item'))}
    tabs['target_user'][5] = tabs['source_user'][5]  # exact ties: the gradient splits 0.5 / 0.5
    dims = [2 * dim] + list(hidden)
    ws = [torch.randn(b, a, generator=g) * 0.2 for a, b in zip(dims[:-1], dims[1:])]
    bs = [torch.randn(b, generator=g) * 0.1 for b in dims[1:]]
    ow, ob = torch.randn(1, dims[-1], generator=g) * 0.3, torch.randn(1, generator=g) * 0.1
    u, i = rand_ids(batch, 150, 220, 1.3), rand_ids(batch, 150, 221)
    u[:3] = 5
    label = (torch.rand(batch, generator=g) < 0.5).float()
    lt = {k: v.clone().requires_grad_(True) for k, v in tabs.items()}
    lw = [w.clone().requires_grad_(True) for w in ws + [ow]]
    lb = [b.clone().requires_grad_(True) for b in bs + [ob]]
    prob = O.dtcdr_neumf_forward(lt, u, i, lw[:-1], lb[:-1], lw[-1], lb[-1])
    ref = O.bce_loss(prob, label)
    (ref * 0.7).backward()
    return tabs, ws + [ow], bs + [ob], u, i, label, lt, lw, lb, ref, prob


@pytest.mark.parametrize('impl,batch,tile_rows', [(0, 70, 0), (1, 1, 0), (1, 70, 32), (1, 129, 64), (1, 300, 0)])
def test_dtcdr_term_matches_oracle(impl, batch, tile_rows):
    tabs, ws, bs, u, i, label, lt, lw, lb, ref, prob = dtcdr_case(batch)
    emu_util.config(sms=3, seed=0)
    r = emu_util.mlp_step(impl, [128, 32, 16, 1], [w.numpy() for w in ws], [b.numpy() for b in bs], ACT_RELU, 1, 1,
                          (tabs['source_user'].numpy(), tabs['target_user'].numpy(), tabs['source_item'].numpy(),
                           tabs['target_item'].numpy(), None), u.numpy(), i.numpy(), label.numpy(), grad_loss=0.7,
                          tile_rows=tile_rows)
    assert abs(r['loss'] - ref.item()) <= LOSS_RTOL * abs(ref.item())
    np.testing.assert_allclose(r['prob'], prob.detach().numpy(), rtol=1e-4, atol=1e-6)
    for k, name in enumerate(('source_user', 'target_user', 'source_item', 'target_item')):
        close(r['dtabs'][k], lt[name].grad, name)
    for l in range(3):
        close(r['dW'][l], lw[l].grad, f'W{l}')
        close(r['db'][l], lb[l].grad, f'b{l}')


def test_tc_supported_stacks():
    L = emu_util.lib()
    import ctypes

    def ok(dims):
        return bool(L.emu_tc_mlp_supported(len(dims) - 1, (ctypes.c_int * len(dims))(*dims)))

    assert ok([64, 128, 64]) and ok([64, 64]) and ok([128, 32, 16, 1]) and ok([128, 64, 32, 1])
    assert not ok([512, 64, 1])          # wider than 256
    assert not ok([64, 12, 64])          # hidden width not a multiple of 8
    assert not ok([256, 64, 32, 1])      # 128 weight-gradient tiles in layer 0: more than a warp carries (8 x 8)
