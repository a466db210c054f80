"""CPU twins of tests/test_gpu_graph.py (BiTGCF graph kernels: CSR SpMM with split rows, propagate, transfer + normalise,
all with backward) -- the same test bodies with the kernels' sources running under the CTA emulator."""
import pytest
import torch

import emu_util
import test_gpu_graph as G


@pytest.fixture(autouse=True)
def on_cpu(monkeypatch):
    """`.cuda()` / device='cuda' in the GPU test bodies become no-ops on CPU tensors for the duration of a test."""
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self.detach().clone().requires_grad_(self.requires_grad)
                        if self.is_leaf else self, raising=False)
    import recbole_cdr_b200.graph as graph
    real = graph.NormAdj
    monkeypatch.setattr(graph, 'NormAdj', lambda r, c, nu, ni, device, **k: real(r, c, nu, ni, 'cpu', **k))
    with emu_util.patched_ops(sms=3):
        yield


@pytest.mark.parametrize('chunk', [4, 256])
@pytest.mark.parametrize('dim,zipf', [(64, None), (32, 1.2), (96, 1.5)])
def test_norm_adj_and_spmm_match_oracle(chunk, dim, zipf):
    G.test_norm_adj_and_spmm_match_oracle(chunk, dim, zipf)


def test_graph_prop_forward_backward():
    G.test_graph_prop_forward_backward()


@pytest.mark.parametrize('dim', [32, 64])
def test_transfer_norm_forward_backward(dim):
    G.test_transfer_norm_forward_backward(dim)
