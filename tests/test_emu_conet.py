"""tc_conet_kernel (one CoNet tower pass + backward + scatter in ONE kernel, 3xTF32 mma.sync tiles) under the CPU CTA
emulator, against the oracle restatement of conet.py:105-197.  Kernel *logic* only; hardware parity is the job of the
``gpu`` tests (tests/test_gpu_engines.py)."""
import numpy as np
import pytest
import torch

import emu_util
from oracle import cdr_oracle as O


def rand_table(n, d, seed, std=0.3):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, d, generator=g) * std


def make_case(batch, dim, hidden, want, overlap_users=True, seed=5):
    g = torch.Generator().manual_seed(seed)
    n_u, n_i, n_ov = 90, 70, 40
    tabs = {k: rand_table(n, dim, 300 + j) for j, (k, n) in enumerate((('source_user', n_u), ('source_item', n_i),
                                                                       ('target_user', n_u), ('target_item', n_i)))}
    dims = [2 * dim] + list(hidden)
    mk = lambda a, b, s=0.25: torch.randn(b, a, generator=g) * s
    P = dict(ws=[mk(a, b) for a, b in zip(dims[:-1], dims[1:])], wt=[mk(a, b) for a, b in zip(dims[:-1], dims[1:])],
             h=[mk(a, b) for a, b in zip(dims[:-1], dims[1:])],
             bs=[torch.randn(b, generator=g) * 0.1 for b in dims[1:]], bt=[torch.randn(b, generator=g) * 0.1 for b in dims[1:]],
             out_s_w=mk(dims[-1], 1, 0.5), out_s_b=torch.randn(1, generator=g) * 0.1,
             out_t_w=mk(dims[-1], 1, 0.5), out_t_b=torch.randn(1, generator=g) * 0.1)
    user = torch.randint(0, n_u, (batch,), generator=g)
    item = torch.randint(0, n_i, (batch,), generator=g)
    user[:4] = 3  # duplicates
    label = (torch.rand(batch, generator=g) < 0.5).float()
    # oracle with autograd
    lt = {k: v.clone().requires_grad_(True) for k, v in tabs.items()}
    lp = {k: ([x.clone().requires_grad_(True) for x in v] if isinstance(v, list) else v.clone().requires_grad_(True))
          for k, v in P.items()}
    ps, pt = O.conet_towers(lt, user, item, lp, overlap_users, n_ov)
    prob = ps if want == 0 else pt
    ref = O.bce_loss(prob, label)
    (ref * 0.9).backward()
    return dims, tabs, P, user, item, label, n_ov, lt, lp, ref, prob


def close(got, want, nm):
    want = np.zeros_like(got) if want is None else want.detach().numpy()
    atol = max(1e-7, 1e-4 * float(np.abs(want).max()))
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=atol, err_msg=nm)


@pytest.mark.parametrize('batch,dim,hidden,want,overlap_users', [
    (70, 32, [32, 16, 8], 0, True),       # the golden CoNet shape: one K chunk, three cross layers
    (70, 32, [32, 16, 8], 1, False),      # target tower, item-overlap mask
    (150, 64, [64, 32, 16, 8], 0, True),  # CoNet.yaml layer stack, two K chunks, ragged last tile, 2 tiles on one CTA
    (64, 128, [64, 32, 16, 8], 1, True),  # config #3 width: four K chunks
    (33, 32, [16], 0, True),              # a single cross layer: the head sits directly on the K-chunked layer
    (1, 32, [32, 8], 1, True),
])
def test_conet_pass_matches_oracle(batch, dim, hidden, want, overlap_users):
    dims, tabs, P, user, item, label, n_ov, lt, lp, ref, prob = make_case(batch, dim, hidden, want, overlap_users)
    emu_util.config(sms=2, seed=0)
    sfx = 's' if want == 0 else 't'
    Pk = dict(ws=P['ws'], bs=P['bs'], wt=P['wt'], bt=P['bt'], h=P['h'], out_w=P[f'out_{sfx}_w'], out_b=P[f'out_{sfx}_b'])
    r = emu_util.conet_step(dims, {k: ([x.numpy() for x in v] if isinstance(v, list) else v.numpy()) for k, v in Pk.items()},
                            want, (tabs['source_user'].numpy(), tabs['source_item'].numpy(), tabs['target_user'].numpy(),
                                   tabs['target_item'].numpy()), user.numpy(), item.numpy(), label.numpy(),
                            mask_on_item=not overlap_users, n_overlap=n_ov, grad_loss=0.9)
    assert abs(r['loss'] - ref.item()) <= 1e-4 * abs(ref.item())
    np.testing.assert_allclose(r['prob'], prob.detach().numpy(), rtol=1e-4, atol=1e-6)
    for k, name in enumerate(('source_user', 'source_item', 'target_user', 'target_item')):
        close(r['dtabs'][k], lt[name].grad, name)
    for key in ('ws', 'bs', 'wt', 'bt', 'h'):
        for l in range(len(hidden)):
            close(r['grads'][key][l], lp[key][l].grad, f'{key}[{l}]')
    close(r['dout_w'], lp[f'out_{sfx}_w'].grad, 'out_w')
    close(r['dout_b'], lp[f'out_{sfx}_b'].grad, 'out_b')


def test_conet_pass_is_schedule_independent():
    dims, tabs, P, user, item, label, n_ov, lt, lp, ref, prob = make_case(100, 32, [32, 16, 8], 0)
    Pk = dict(ws=P['ws'], bs=P['bs'], wt=P['wt'], bt=P['bt'], h=P['h'], out_w=P['out_s_w'], out_b=P['out_s_b'])
    Pk = {k: ([x.numpy() for x in v] if isinstance(v, list) else v.numpy()) for k, v in Pk.items()}
    outs = []
    for seed in (0, 3, 11):
        emu_util.config(sms=1, seed=seed)
        outs.append(emu_util.conet_step(dims, Pk, 0, (tabs['source_user'].numpy(), tabs['source_item'].numpy(),
                                                      tabs['target_user'].numpy(), tabs['target_item'].numpy()),
                                        user.numpy(), item.numpy(), label.numpy(), mask_on_item=False, n_overlap=n_ov))
    emu_util.config(sms=4, seed=0)
    for r in outs[1:]:
        assert r['loss'] == outs[0]['loss']
        np.testing.assert_array_equal(r['grads']['h'][0], outs[0]['grads']['h'][0])
        np.testing.assert_allclose(r['dtabs'][0], outs[0]['dtabs'][0], rtol=1e-5, atol=1e-9)


def test_conet_supported_stacks():
    import ctypes
    L = emu_util.lib()

    def ok(dims, dim):
        return bool(L.emu_conet_supported(len(dims) - 1, (ctypes.c_int * len(dims))(*dims), dim))

    assert ok([256, 64, 32, 16, 8], 128) and ok([128, 64, 32, 16, 8], 64) and ok([64, 32, 16, 8], 32)
    assert not ok([32, 16, 8], 16)           # 2*dim must be a multiple of the 64-column K chunk
    assert not ok([256, 128, 64], 128)       # hidden wider than 64
    assert not ok([256, 64, 12], 128)        # hidden width not a multiple of 8


@pytest.mark.parametrize('n_ov_case,batch,sms', [('all', 80, 2), ('none', 80, 2), ('mixed_no_perm', 2200, 1)])
def test_conet_row_reordering_and_cross_term_skipping(n_ov_case, batch, sms):
    """Rows are re-ordered per CTA (overlapped first) so that 16-row MMA tiles without an overlapped row skip their cross
    products: all rows overlapped (nothing skipped), none overlapped (every cross product skipped), and a CTA with more rows
    than the re-ordering buffer holds (natural order, skipping still decided per tile)."""
    dims, tabs, P, user, item, label, n_ov, lt, lp, ref, prob = make_case(batch, 32, [16], 0, True)
    n_ov = {'all': 10 ** 6, 'none': 0, 'mixed_no_perm': n_ov}[n_ov_case]
    lt = {k: v.detach().clone().requires_grad_(True) for k, v in tabs.items()}
    lp = {k: ([x.detach().clone().requires_grad_(True) for x in v] if isinstance(v, list) else v.detach().clone().requires_grad_(True))
          for k, v in P.items()}
    ps, _ = O.conet_towers(lt, user, item, lp, True, n_ov)
    ref = O.bce_loss(ps, label)
    ref.backward()
    emu_util.config(sms=sms, seed=0)
    Pk = dict(ws=P['ws'], bs=P['bs'], wt=P['wt'], bt=P['bt'], h=P['h'], out_w=P['out_s_w'], out_b=P['out_s_b'])
    r = emu_util.conet_step(dims, {k: ([x.numpy() for x in v] if isinstance(v, list) else v.numpy()) for k, v in Pk.items()}, 0,
                            (tabs['source_user'].numpy(), tabs['source_item'].numpy(), tabs['target_user'].numpy(),
                             tabs['target_item'].numpy()), user.numpy(), item.numpy(), label.numpy(), mask_on_item=False,
                            n_overlap=n_ov)
    assert abs(r['loss'] - ref.item()) <= 1e-4 * abs(ref.item())
    np.testing.assert_allclose(r['prob'], ps.detach().numpy(), rtol=1e-4, atol=1e-6)
    for k, name in enumerate(('source_user', 'source_item', 'target_user', 'target_item')):
        close(r['dtabs'][k], lt[name].grad, name)
    for key in ('ws', 'bs', 'wt', 'bt', 'h'):
        close(r['grads'][key][0], lp[key][0].grad, f'{key}[0]')
