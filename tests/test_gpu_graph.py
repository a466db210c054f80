"""GPU parity of the BiTGCF graph kernels (A9-A12) and of the BiTGCF drop-in class against the oracle / goldens."""
import numpy as np
import pytest
import torch

from oracle import cdr_oracle as O
from fake_data import FakeDataset, base_config
from golden_util import Golden, bitgcf_graph

pytestmark = pytest.mark.gpu


def random_graph(n_users, n_items, n_edges, seed, zipf=None):
    rng = np.random.RandomState(seed)
    r = rng.randint(0, n_users, n_edges)
    c = np.minimum(rng.zipf(zipf, n_edges) - 1, n_items - 1) if zipf else rng.randint(0, n_items, n_edges)
    return r, c


@pytest.mark.parametrize('chunk', [4, 256])
@pytest.mark.parametrize('dim,zipf', [(64, None), (64, 1.2), (32, None), (96, 1.5), (128, None)])
def test_norm_adj_and_spmm_match_oracle(chunk, dim, zipf):
    from recbole_cdr_b200.graph import NormAdj
    nu, ni = 300, 200
    r, c = random_graph(nu, ni, 4000, 1, zipf)
    adj = NormAdj(r, c, nu, ni, 'cuda', chunk=chunk)
    ref = O.bitgcf_norm_adj(r, c, nu, ni)
    got = adj.to_sparse_coo().cpu()
    assert torch.equal(got.indices(), ref.indices()) and torch.equal(got.values(), ref.values())   # A9: bit-exact
    X = torch.randn(nu + ni, dim, generator=torch.Generator().manual_seed(2))
    S = adj.spmm(X.cuda())
    torch.testing.assert_close(S.cpu(), torch.sparse.mm(ref, X), rtol=1e-4, atol=1e-5)
    if chunk == 4:
        assert adj.split_rows.numel() > 0           # the split-row (atomic) path is exercised


def test_graph_prop_forward_backward():
    from recbole_cdr_b200.graph import GraphProp, NormAdj
    nu, ni, dim = 150, 120, 64
    r, c = random_graph(nu, ni, 2500, 3, 1.3)
    adj = NormAdj(r, c, nu, ni, 'cuda', chunk=16)
    ref_adj = O.bitgcf_norm_adj(r, c, nu, ni)
    g = torch.Generator().manual_seed(4)
    E, G = torch.randn(nu + ni, dim, generator=g), torch.randn(nu + ni, dim, generator=g)
    er = E.clone().requires_grad_(True)
    yr = O.bitgcf_graph_layer(ref_adj, er)
    (yr * G).sum().backward()
    ec = E.cuda().requires_grad_(True)
    yc = GraphProp.apply(ec, adj)
    torch.testing.assert_close(yc.cpu(), yr.detach(), rtol=1e-4, atol=1e-5)
    (yc * G.cuda()).sum().backward()
    torch.testing.assert_close(ec.grad.cpu(), er.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('dim', [32, 64])
def test_transfer_norm_forward_backward(dim):
    from recbole_cdr_b200.graph import TransferNorm
    nu, ni, ovu, ovi = 90, 70, 31, 17
    g = torch.Generator().manual_seed(5)
    Ps, Pt = torch.randn(nu + ni, dim, generator=g), torch.randn(nu + ni, dim, generator=g)
    Ps[3] = 0.0                                     # a zero row: F.normalize's eps branch
    deg = {k: torch.randint(0, 9, (n, 1), generator=g).float() for k, n in (('su', nu), ('tu', nu), ('si', ni), ('ti', ni))}
    Gs = [torch.randn(nu + ni, dim, generator=g) for _ in range(4)]
    a, b = Ps.clone().requires_grad_(True), Pt.clone().requires_grad_(True)
    es, et = O.bitgcf_transfer_layer(a, b, nu, ni, ovu, ovi, 0.8, 0.7, deg)
    ns, nt = torch.nn.functional.normalize(es, p=2, dim=1), torch.nn.functional.normalize(et, p=2, dim=1)
    (es * Gs[0] + et * Gs[1] + ns * Gs[2] + nt * Gs[3]).sum().backward()
    ac, bc = Ps.cuda().requires_grad_(True), Pt.cuda().requires_grad_(True)
    ds = torch.cat([deg['su'], deg['si']]).reshape(-1).cuda()
    dt = torch.cat([deg['tu'], deg['ti']]).reshape(-1).cuda()
    out = TransferNorm.apply(ac, bc, ds, dt, nu, ni, ovu, ovi, 0.8, 0.7)
    for got, ref in zip(out, (es, et, ns, nt)):
        torch.testing.assert_close(got.cpu(), ref.detach(), rtol=1e-4, atol=1e-6)
    sum((o * G.cuda()).sum() for o, G in zip(out, Gs)).backward()
    keep = torch.ones(nu + ni, dtype=torch.bool)
    keep[3] = False                                 # torch's normalize backward at an exactly-zero row is 0/eps noise
    torch.testing.assert_close(ac.grad.cpu()[keep], a.grad[keep], rtol=2e-4, atol=1e-5)
    torch.testing.assert_close(bc.grad.cpu()[keep], b.grad[keep], rtol=2e-4, atol=1e-5)


@pytest.mark.parametrize('way', ['concat', 'mean'])
def test_bitgcf_model_matches_reference_golden(way):
    from recbole_cdr_b200.data import Interaction
    from recbole_cdr_b200.model.cross_domain_recommender.bitgcf import BiTGCF
    g = Golden(f'bitgcf_{way}')
    _, _, edges, _ = bitgcf_graph(g)
    ds = FakeDataset.from_golden(g, edges)
    m = BiTGCF(base_config(embedding_size=32, n_layers=2, reg_weight=0.001, lambda_source=0.8, lambda_target=0.7,
                           drop_rate=0.0, connect_way=way), ds)
    m.load_state_dict({n: g.param(n) for n in g.param_names()}, strict=True)
    m = m.to('cuda')
    batch = Interaction({k[len('batch/'):]: torch.from_numpy(g.z[k]) for k in g.z.files if k.startswith('batch/')}).to('cuda')
    losses = m.calculate_loss(batch)
    assert isinstance(losses, tuple) and len(losses) == 2 and all(l.shape == (1,) for l in losses)
    for got, ref in zip(losses, g.losses()):
        torch.testing.assert_close(got.detach().cpu(), ref, rtol=1e-4, atol=0)
    sum(losses).sum().backward()
    for name, p in m.named_parameters():
        torch.testing.assert_close(p.grad.cpu(), g.grad(name), rtol=2e-4, atol=2e-7, msg=lambda s: f'grad {name}: {s}')
    torch.testing.assert_close(m.predict(batch).cpu(), g.t('predict'), rtol=1e-4, atol=1e-6)
    full = m.full_sort_predict(Interaction({'target_user_id': batch['target_user_id'][:5]})).view(5, -1)
    assert full.shape[1] == m.target_num_items
