"""CPU twins of tests/test_gpu_kernels.py: the SAME test bodies, with ``dev()`` pointing at the CPU and
``recbole_cdr_b200.ops`` running the kernels' sources under the CTA emulator (small parametrisations only -- the emulator
is slow).  Edge cases covered: empty and single-row batches, the [b, 1] overlap batch, concat buffers, Zipf and
all-duplicate ids, out-of-range ids (IndexError like nn.Embedding), saturated logits, in-place gradient mode, the fused
SGD scatter, the cross-stitch dense unit."""
import pytest
import torch

import emu_util
import test_gpu_kernels as G


@pytest.fixture(autouse=True)
def on_cpu(monkeypatch):
    monkeypatch.setattr(G, 'dev', lambda: torch.device('cpu'))
    with emu_util.patched_ops(sms=3):
        yield


@pytest.mark.parametrize('dim,n_idx', [(4, 0), (64, 1), (96, 5), (256, 300)])
def test_gather_rows_bit_exact(dim, n_idx):
    G.test_gather_rows_bit_exact(dim, n_idx)


def test_gather_rows_into_concat_buffer_and_2d_index():
    G.test_gather_rows_into_concat_buffer_and_2d_index()


@pytest.mark.parametrize('dim,zipf', [(64, 1.05), (36, None)])
def test_scatter_add_matches_index_add(dim, zipf):
    G.test_scatter_add_matches_index_add(dim, zipf)


def test_scatter_add_all_duplicates():
    G.test_scatter_add_all_duplicates()


def test_gather_autograd_backward_is_dense_index_add():
    G.test_gather_autograd_backward_is_dense_index_add()


def test_out_of_range_id_raises_index_error():
    G.test_out_of_range_id_raises_index_error()


@pytest.mark.parametrize('batch,dim,nu,ni,zipf', [(1, 64, 50, 60, None), (3, 64, 50, 60, None), (257, 64, 1000, 1200, None),
                                                  (1000, 32, 300, 300, 1.3), (513, 96, 300, 300, None), (100, 36, 40, 40, None)])
def test_bpr_loss_and_grads(batch, dim, nu, ni, zipf):
    G.test_bpr_loss_and_grads(batch, dim, nu, ni, zipf)


@pytest.mark.parametrize('kind', ['mse', 'bce', 'none'])
@pytest.mark.parametrize('batch,dim,zipf', [(5, 64, None), (300, 96, None), (600, 64, 1.05)])
def test_point_loss_and_grads(kind, batch, dim, zipf):
    G.test_point_loss_and_grads(kind, batch, dim, zipf)


def test_inplace_table_grad_mode_matches_autograd_mode(monkeypatch):
    # on the CPU `tensor.to(dev())` returns the tensor itself, so the two modes of the GPU test body would share one leaf
    # (and its .grad); hand out fresh copies instead
    real = G.rand_table
    monkeypatch.setattr(G, 'rand_table', lambda *a, **k: _Fresh(real(*a, **k)))
    G.test_inplace_table_grad_mode_matches_autograd_mode()


class _Fresh:
    """Stands in for a tensor whose ``.to(device)`` always yields a new leaf (what a host->device copy does on the GPU)."""

    def __init__(self, t):
        self.t = t

    def to(self, *_a, **_k):
        return self.t.clone()


def test_fused_sgd_scatter_updates_weights_in_place(monkeypatch):
    real = G.rand_table
    monkeypatch.setattr(G, 'rand_table', lambda *a, **k: _Aliasless(real(*a, **k)))
    G.test_fused_sgd_scatter_updates_weights_in_place()


class _Aliasless(torch.Tensor):
    """A tensor whose ``.to(device)`` copies even when the device does not change (what host->device does on the GPU), so
    that the kernel's in-place update does not reach the host-side original the test computes its reference from."""

    @staticmethod
    def __new__(cls, t):
        return torch.Tensor._make_subclass(cls, t)

    def to(self, *a, **k):
        return self.as_subclass(torch.Tensor).clone()


@pytest.mark.parametrize('M,N,K', [(1, 1, 8), (100, 64, 128), (777, 33, 20), (300, 24, 32), (257, 16, 64), (130, 128, 36)])
@pytest.mark.parametrize('act', ['none', 'relu', 'tanh', 'sigmoid'])
def test_dense_fwd_bwd(M, N, K, act):
    G.test_dense_fwd_bwd(M, N, K, act)


def test_mse_rows_and_bce_logit():
    G.test_mse_rows_and_bce_logit()


@pytest.mark.parametrize('shapes', [[(64, 256), (32, 64), (16, 32), (8, 16)], [(1, 1)], [(5, 3), (7, 1), (30, 33)], [(2, 2)] * 8])
def test_frob_sum(shapes):
    G.test_frob_sum(shapes)


def test_frob_sum_refuses_more_matrices_than_the_launch_carries():
    from recbole_cdr_b200 import _lib, ops
    with pytest.raises(_lib.XdrError, match='n_mats'):
        ops.frob_sum([torch.ones(2, 2) for _ in range(9)])


def test_gather_max2_concat_fwd_bwd():
    G.test_gather_max2_concat_fwd_bwd()


def test_frozen_tables_get_no_gradient_and_no_dense_allocation():
    """the row gathers consult needs_input_grad -- a table with requires_grad = False is neither scattered into nor
    given a dense [N, D] zero gradient."""
    with emu_util.patched_ops() as ops:
        g = torch.Generator().manual_seed(5)
        ut, it = torch.randn(50, 16, generator=g), torch.randn(60, 16, generator=g)
        u, i = torch.randint(0, 50, (33,), generator=g), torch.randint(0, 60, (33,), generator=g)
        ut.requires_grad_(False)
        it.requires_grad_(True)
        seen, real = [], ops.call
        ops.call = lambda nm, *a, **k: (seen.append(nm), real(nm, *a, **k))[1]
        try:
            out = ops.gather_concat(ut, it, u, i) if hasattr(ops, 'gather_concat') else ops.GatherConcat.apply(ut, it, u, i)
            out.sum().backward()
            rows = ops.gather_rows(ut, u)          # nothing requires grad here: no backward at all
            assert not rows.requires_grad
        finally:
            ops.call = real
        assert ut.grad is None and it.grad is not None
        assert seen.count('xdr_scatter_add_rows') == 1          # the item table only
        ref = torch.zeros_like(it).index_add_(0, i, torch.ones(33, 16))
        torch.testing.assert_close(it.grad, ref)
