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


def test_run_lanes_forks_after_main_and_joins_every_side_stream(monkeypatch):
    """ops._run_lanes with stand-in streams: every side stream first waits for the caller's stream (everything queued so far),
    lanes i with i % n == k run on side stream k - 1 (k = 0: the caller's stream), and the caller's stream waits for every side
    stream that was used before anything after the node can run -- the fork / join shape a CUDA-graph capture records."""
    import contextlib
    from recbole_cdr_b200 import ops
    log = []

    class FakeStream:
        def __init__(self, name):
            self.name = name

        def wait_stream(self, other):
            log.append(('wait', self.name, other.name))

    main = FakeStream('main')
    sides = [FakeStream(f'side{i}') for i in range(3)]
    current = [main]

    @contextlib.contextmanager
    def fake_stream_ctx(st):
        current.append(st)
        try:
            yield
        finally:
            current.pop()

    monkeypatch.setattr(torch.cuda, 'current_stream', lambda device=None: current[-1])
    monkeypatch.setattr(torch.cuda, 'stream', fake_stream_ctx)
    monkeypatch.setattr(torch.cuda, 'is_current_stream_capturing', lambda: False)
    monkeypatch.setattr(ops, 'side_streams', lambda device, n: sides[:n])
    dev = torch.device('cuda', 0)
    lanes = [(lambda i=i: log.append(('lane', i, current[-1].name))) for i in range(4)]

    prev = ops.set_cross_streams(2)
    try:
        ops._run_lanes(dev, lanes)
        assert log == [('wait', 'side0', 'main'), ('lane', 1, 'side0'), ('lane', 3, 'side0'), ('lane', 0, 'main'),
                       ('lane', 2, 'main'), ('wait', 'main', 'side0')]
        log.clear()
        ops.set_cross_streams(4)
        ops._run_lanes(dev, lanes[:3])           # fewer lanes than streams: one stream per lane
        assert log == [('wait', 'side0', 'main'), ('lane', 1, 'side0'), ('wait', 'side1', 'main'), ('lane', 2, 'side1'),
                       ('lane', 0, 'main'), ('wait', 'main', 'side0'), ('wait', 'main', 'side1')]
        log.clear()
        ops.set_cross_streams(-1)                # default: lanes only inside a capture
        ops._run_lanes(dev, lanes)
        assert [e for e in log if e[0] == 'wait'] == [] and [e[2] for e in log] == ['main'] * 4
        log.clear()
        monkeypatch.setattr(torch.cuda, 'is_current_stream_capturing', lambda: True)
        ops._run_lanes(dev, lanes)
        assert log[0] == ('wait', 'side0', 'main') and log[-1] == ('wait', 'main', 'side0')
        assert sorted(e[1] for e in log if e[0] == 'lane' and e[2] == 'side0') == [1, 3]
    finally:
        ops.set_cross_streams(prev)
