"""GPU parity tests of kernels written while no GPU was reachable (marker ``unvalidated``: skipped unless
XDR_RUN_UNVALIDATED=1).  Their logic already runs under the CPU CTA emulator (tests/test_emu_*.py); these are the hardware
counterparts, at the tolerances of tests/test_gpu_kernels.py / test_gpu_models.py.  Once a test here has passed on a B200 it
moves to the regular ``gpu`` files."""
import numpy as np
import pytest
import torch

from fake_data import FakeDataset, base_config  # noqa: F401
from golden_util import Golden
from oracle import cdr_oracle as O
from test_gpu_kernels import LOSS_RTOL, dev, lib, ops, rand_ids, rand_table
from test_gpu_models import EMCDR_CFG, build, check_loss_and_grads, cuda_batch

pytestmark = [pytest.mark.gpu, pytest.mark.unvalidated]


# ------------------------------------------------------------------------------------------ tensor-core fused MLP (tc_mlp.cu)
@pytest.mark.parametrize('batch', [1, 31, 32, 33, 63, 64, 65, 1000, 8192, 20000])
def test_tc_mlp_map_loss_matches_oracle(batch):
    """EMCDR map step through the 3xTF32 tile kernel at sizes around the 32- and 64-row tiles."""
    g = torch.Generator().manual_seed(111)
    src, tgt = rand_table(3000, 64, 112, 0.3), rand_table(3000, 64, 113, 0.3)
    ws = [torch.randn(128, 64, generator=g) * 0.2, torch.randn(64, 128, generator=g) * 0.2]
    bs = [torch.randn(128, generator=g) * 0.1, torch.randn(64, generator=g) * 0.1]
    idx = rand_ids(batch, 3000, 114, 1.3)
    leaves = [t.clone().requires_grad_(True) for t in [src, tgt] + ws + bs]
    ref = O.emcdr_map_loss(leaves[0], leaves[1], idx.view(-1, 1), leaves[2:4], leaves[4:6])
    ref.backward()
    c = [t.to(dev()).requires_grad_(True) for t in [src, tgt] + ws + bs]
    assert ops().fused_mlp_supported([64, 128, 64], 'tc')
    loss = ops().fused_mlp_loss(0, 0, lib().ACT_TANH, idx.to(dev()), None, None, (c[0], None, None, None, c[1]), c[2:4], c[4:6],
                                'tc')
    torch.testing.assert_close(loss.cpu(), ref.detach(), rtol=LOSS_RTOL, atol=0)
    (loss * 1.3).backward()
    for got, want, nm in zip(c, leaves, ('src', 'tgt', 'W1', 'W2', 'b1', 'b2')):
        atol = max(1e-7, 1e-4 * want.grad.abs().max().item())
        torch.testing.assert_close(got.grad.cpu(), want.grad * 1.3, rtol=2e-4, atol=atol, msg=lambda s: f'{nm}: {s}')


def test_tc_mlp_supported_stacks():
    assert ops().fused_mlp_supported([64, 128, 64], 'tc') and ops().fused_mlp_supported([128, 32, 16, 1], 'tc')
    assert not ops().fused_mlp_supported([512, 64, 1], 'tc')
    assert not ops().fused_mlp_supported([64, 12, 64], 'tc')


@pytest.mark.parametrize('case,mf', [('non_linear', 'non_linear'), ('linear', 'linear'), ('items', 'non_linear')])
def test_emcdr_map_phase_tc_engine(case, mf):
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    g = Golden(f'emcdr_map_{case}')
    m = build(EMCDR, g, dict(EMCDR_CFG, latent_factor_model='BPR', mapping_function=mf, xdr_fused_mlp='tc'))
    assert m.fused_mlp_engine == 'tc'
    m.set_phase('OVERLAP')
    check_loss_and_grads(m, g, cuda_batch(g), grad_atol=1e-6)


def test_dtcdr_tc_engine():
    from recbole_cdr_b200.model.cross_domain_recommender.dtcdr import DTCDR
    g = Golden('dtcdr_neumf')
    m = build(DTCDR, g, dict(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF',
                             alpha=g.meta('alpha'), xdr_fused_mlp='tc'))
    assert m._fused_ok() and m.fused_mlp_engine == 'tc'
    batch = cuda_batch(g)
    check_loss_and_grads(m, g, batch, grad_rtol=2e-4, grad_atol=2e-6)
    torch.testing.assert_close(m.predict(batch).cpu(), g.t('predict'), rtol=1e-4, atol=1e-6)
