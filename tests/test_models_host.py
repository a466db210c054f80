"""CPU checks of the drop-in model classes (no kernel is launched): construction from the reference's (config, dataset)
contract, state_dict keys and shapes equal to the reference's, and -- same construction and init order -- bit-identical
seeded initial weights (the goldens were produced by the unmodified reference classes under torch.manual_seed(2022))."""
import pytest
import torch

from fake_data import FakeDataset, base_config
from golden_util import Golden, bitgcf_graph

CASES = {
    'emcdr_bpr_source': ('EMCDR', dict(source_embedding_size=64, target_embedding_size=64, reg_weight=0.01,
                                       mlp_hidden_size=[128], latent_factor_model='BPR', mapping_function='non_linear')),
    'emcdr_map_linear': ('EMCDR', dict(source_embedding_size=64, target_embedding_size=64, reg_weight=0.01,
                                       mlp_hidden_size=[128], latent_factor_model='BPR', mapping_function='linear')),
    'cmf_both': ('CMF', {'embedding_size': 64, 'alpha': 0.3, 'lambda': 0.05, 'gamma': 0.02}),
    'conet_users': ('CoNet', dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8])),
    'conet_items': ('CoNet', dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8])),
    'dtcdr_neumf': ('DTCDR', dict(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF', alpha=0.4)),
    'bitgcf_concat': ('BiTGCF', dict(embedding_size=32, n_layers=2, reg_weight=0.001, lambda_source=0.8, lambda_target=0.7,
                                     drop_rate=0.0, connect_way='concat')),
}


@pytest.mark.parametrize('golden', sorted(CASES))
def test_seeded_construction_matches_reference(golden):
    from recbole_cdr_b200.utils import get_model, ModelType
    name, cfg = CASES[golden]
    g = Golden(golden)
    edges = bitgcf_graph(g)[2] if name == 'BiTGCF' else None
    torch.manual_seed(2022)
    m = get_model(name)(base_config(device='cpu', **cfg), FakeDataset.from_golden(g, edges))
    assert m.type == ModelType.CROSSDOMAIN
    params = dict(m.named_parameters())
    assert set(params) == set(g.param_names())                     # the reference's state_dict keys, nothing else
    assert set(m.state_dict()) == set(g.param_names())
    for n, p in params.items():
        assert torch.equal(p.detach(), g.param(n)), n              # bit-identical seeded init


def test_emcdr_phase_dispatch_and_mode_rules():
    from recbole_cdr_b200.model.cross_domain_recommender.emcdr import EMCDR
    from recbole_cdr_b200.utils import InputType
    g = Golden('emcdr_bpr_source')
    cfg = base_config(device='cpu', **CASES['emcdr_bpr_source'][1])
    m = EMCDR(cfg, FakeDataset.from_golden(g))
    assert m.mode == 'overlap_users' and m.phase == 'both' and m.input_type == InputType.PAIRWISE
    m.set_phase('SOURCE')
    assert m.fused_step_spec()['fields'] == ['source_user_id', 'source_item_id', 'neg_source_item_id']
    m.set_phase('BOTH')                                           # anything but SOURCE/OVERLAP is the target loss (emcdr.py:170-176)
    assert m.fused_step_spec()['fields'][0] == 'target_user_id'
    m.set_phase('OVERLAP')
    assert m.fused_step_spec() is None
    with pytest.raises(AssertionError):                            # both users and items overlapped: EMCDR refuses (emcdr.py:33-34)
        EMCDR(cfg, FakeDataset(5, 3, 3, 5, 3, 3))
    mf = EMCDR(base_config(device='cpu', **dict(CASES['emcdr_bpr_source'][1], latent_factor_model='MF')), FakeDataset.from_golden(g))
    assert mf.input_type == InputType.POINTWISE and mf.SOURCE_LABEL == 'source_label'


def test_dtcdr_dmf_is_out_of_scope():
    from recbole_cdr_b200.model.cross_domain_recommender.dtcdr import DTCDR
    g = Golden('dtcdr_neumf')
    with pytest.raises(NotImplementedError):
        DTCDR(base_config(device='cpu', **dict(CASES['dtcdr_neumf'][1], base_model='DMF')), FakeDataset.from_golden(g))


def test_conet_hot_path_switches(monkeypatch):
    """CoNet's host-side choices: one stacked pass per BOTH step and the tcgen05 dense engine by default; config keys win over
    the environment (``XDR_CONET_STACK``, ``XDR_DENSE_ENGINE``), the environment over the defaults."""
    from recbole_cdr_b200.model.cross_domain_recommender.conet import CoNet
    ds = FakeDataset(41, 30, 35, 1, 50, 60)
    mk = lambda **kw: CoNet(base_config(device='cpu', embedding_size=16, reg_weight=0.0, mlp_hidden_size=[16, 8], **kw), ds)
    monkeypatch.delenv('XDR_CONET_STACK', raising=False)
    monkeypatch.delenv('XDR_DENSE_ENGINE', raising=False)
    m = mk()
    assert m.stack_passes and m.dense_engine == 1 and not m.use_fused_conet
    m = mk(xdr_stack_passes=False, xdr_dense_engine=0)
    assert not m.stack_passes and m.dense_engine == 0
    monkeypatch.setenv('XDR_CONET_STACK', '0')
    monkeypatch.setenv('XDR_DENSE_ENGINE', '2')
    m = mk()
    assert not m.stack_passes and m.dense_engine == 2
    m = mk(xdr_stack_passes=True, xdr_dense_engine=1)
    assert m.stack_passes and m.dense_engine == 1


def test_cross_streams_setting_round_trip():
    from recbole_cdr_b200 import ops
    prev = ops.set_cross_streams(4)
    try:
        assert ops.set_cross_streams(-5) == 4      # clamped to -1 = "two lanes inside a graph capture, none in eager steps"
        assert ops.set_cross_streams(0) == -1
        ran = []
        ops._run_lanes(torch.device('cpu'), [lambda: ran.append(0), lambda: ran.append(1), lambda: ran.append(2)])
        assert ran == [0, 1, 2]                    # no CUDA device: every lane in order on the caller's stream
    finally:
        ops.set_cross_streams(prev)
