"""GPU: the trainer mirror end to end (the reference's own test style, tests/test_model.py:10-11 -- one epoch per
phase must run -- plus what the reference never asserts: the loss goes down, and the fused multi-step path follows the
batch-by-batch SGD path)."""
import numpy as np
import pytest
import torch

from fake_data import FakeDataset, base_config

pytestmark = pytest.mark.gpu


def make_world(pairwise, seed=0, device='cuda'):
    from recbole_cdr_b200.data import CrossDomainDataloader, DomainTrainDataLoader, OverlapDataloader
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler, TargetDomainSampler
    ds = FakeDataset(201, 300, 280, 1, 500, 450)        # user-overlap layout (EMCDR / CoNet)
    rng = np.random.RandomState(seed)
    su, si = ds.valid_ids('source')
    tu, ti = ds.valid_ids('target')
    s_u, s_i = rng.choice(su, 6000), rng.choice(si, 6000)
    t_u, t_i = rng.choice(tu, 5000), rng.choice(ti, 5000)
    s_smp = CrossDomainSourceSampler('train', ds, user_ids=s_u, item_ids=s_i, device=device).set_phase('train')
    t_smp = TargetDomainSampler(ds.num_total_user, ds.target_domain_dataset.num('target_item_id'), t_u, t_i, device=device)
    g = torch.Generator().manual_seed(seed)
    src = DomainTrainDataLoader('source_user_id', 'source_item_id', s_u, s_i, 1024, s_smp, pairwise, 'source_label',
                                shuffle=True, generator=g)
    tgt = DomainTrainDataLoader('target_user_id', 'target_item_id', t_u, t_i, 1024, t_smp, pairwise, 'target_label',
                                shuffle=True, generator=g)
    return ds, CrossDomainDataloader(src, tgt, OverlapDataloader(ds.num_overlap_user, 100, True, g))


def emcdr_cfg(**kw):
    cfg = base_config(latent_factor_model='BPR', source_embedding_size=64, target_embedding_size=64, reg_weight=0.01,
                      mapping_function='non_linear', mlp_hidden_size=[128], learner='adam', learning_rate=0.01,
                      weight_decay=0.0, train_modes=['SOURCE', 'TARGET', 'OVERLAP'], epoch_num=['2', '2', '2'],
                      source_split=False)
    cfg.update(kw)
    return cfg


def test_emcdr_three_phases_train_and_losses_fall():
    from recbole_cdr_b200.utils import get_model, get_trainer, ModelType
    ds, loader = make_world(pairwise=True)
    cfg = emcdr_cfg()
    torch.manual_seed(2022)
    model = get_model('EMCDR')(cfg, ds).to('cuda')
    trainer = get_trainer(ModelType.CROSSDOMAIN, 'EMCDR')(cfg, model)
    seen = []
    trainer.fit(loader, callback_fn=lambda epoch, loss: seen.append((model.phase, epoch, loss)))
    assert [p for p, _, _ in seen] == ['SOURCE', 'SOURCE', 'TARGET', 'TARGET', 'OVERLAP', 'OVERLAP']
    for phase in ('SOURCE', 'TARGET', 'OVERLAP'):
        l = [x for p, _, x in seen if p == phase]
        assert np.isfinite(l).all() and l[1] < l[0], (phase, l)
    assert model.phase == 'OVERLAP'                     # trainer.py:75
    from recbole_cdr_b200.data import Interaction
    inter = Interaction({'target_user_id': torch.arange(1, 50), 'target_item_id': torch.arange(1, 50)}).to('cuda')
    assert torch.isfinite(model.predict(inter)).all()


@pytest.mark.parametrize('name,cfg', [
    ('CMF', dict(embedding_size=64, alpha=0.5, gamma=0.0, **{'lambda': 0.0})),
    ('CoNet', dict(embedding_size=32, reg_weight=0.01, mlp_hidden_size=[32, 16, 8])),
    ('DTCDR', dict(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF', alpha=0.5))])
def test_pointwise_models_one_both_epoch(name, cfg):
    from recbole_cdr_b200.utils import get_model, get_trainer, ModelType
    ds, loader = make_world(pairwise=False)
    full = base_config(learner='adam', learning_rate=0.01, weight_decay=0.0, train_modes=['BOTH'], epoch_num=['2'],
                       source_split=False, **cfg)
    torch.manual_seed(2022)
    model = get_model(name)(full, ds).to('cuda')
    trainer = get_trainer(ModelType.CROSSDOMAIN, name)(full, model)
    seen = []
    trainer.fit(loader, callback_fn=lambda epoch, loss: seen.append(loss))
    assert len(seen) == 2 and np.isfinite(seen).all() and seen[1] < seen[0]


def test_fused_sgd_epoch_tracks_the_per_batch_sgd_path():
    """xdr_fused_steps: K batches per persistent launch with the SGD update fused.  Within an epoch the batches are
    identical for both paths (same loader seed); the fused path reads rows at most a few steps stale, so epoch losses
    agree closely but not bitwise."""
    from recbole_cdr_b200.utils import get_model, get_trainer, ModelType
    out = {}
    for fused in (0, 8):
        ds, loader = make_world(pairwise=True, seed=3)
        cfg = emcdr_cfg(learner='sgd', learning_rate=20.0, reg_weight=0.0, train_modes=['SOURCE'], epoch_num=['4'],
                        xdr_fused_steps=fused)
        torch.manual_seed(2022)
        model = get_model('EMCDR')(cfg, ds).to('cuda')
        trainer = get_trainer(ModelType.CROSSDOMAIN, 'EMCDR')(cfg, model)
        seen = []
        trainer.fit(loader, callback_fn=lambda epoch, loss: seen.append(loss))
        out[fused] = seen
    assert out[8][-1] < out[8][0] - 1e-3 and out[0][-1] < out[0][0] - 1e-3
    np.testing.assert_allclose(out[8], out[0], rtol=5e-3)


def test_graphed_step_equals_eager_step():
    """CUDA-graph replay of (calculate_loss + backward) gives the eager step's loss and gradients, batch after batch."""
    from recbole_cdr_b200 import ops
    from recbole_cdr_b200.data import Interaction
    from recbole_cdr_b200.trainer import GraphedTrainStep
    from recbole_cdr_b200.utils import get_model
    ds = FakeDataset(201, 300, 280, 101, 500, 450)
    cfg = base_config(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF', alpha=0.4)
    torch.manual_seed(1)
    model = get_model('DTCDR')(cfg, ds).to('cuda')
    rng = np.random.RandomState(0)

    def batch(seed):
        from fake_data import make_batch
        r = np.random.RandomState(seed)
        b = make_batch(ds, 'source', 512, r)
        b.update(make_batch(ds, 'target', 512, r))
        return Interaction(b).to('cuda')

    import copy
    eager = copy.deepcopy(model)
    step = GraphedTrainStep(model, batch(0))
    for seed in (1, 2, 3):
        b = batch(seed)
        step.zero_table_grads()
        loss_g = step(b).clone()
        eager.zero_grad(set_to_none=True)
        loss_e = eager.calculate_loss(b)
        loss_e.sum().backward()
        torch.testing.assert_close(loss_g, loss_e.detach().sum(), rtol=1e-6, atol=0)
        ge = dict(eager.named_parameters())
        for n, p in model.named_parameters():
            torch.testing.assert_close(p.grad, ge[n].grad, rtol=1e-4, atol=1e-7, msg=lambda s: f'{n}: {s}')


@pytest.mark.parametrize('streams', [-1, 0, 4])
def test_graphed_conet_step_equals_eager_step(streams):
    """CoNet's stacked BOTH step (ops.cross_pair on the tcgen05 engine, ops.frob_sum) captured as a CUDA graph against the eager
    step -- with the independent launches of every cross-stitch layer on parallel streams (``ops.set_cross_streams``: forked
    and joined inside the autograd nodes, parallel branches of the captured graph; -1, the default: two lanes inside a capture,
    none in eager steps), on one stream (0), and on four lanes everywhere (4)."""
    from recbole_cdr_b200 import ops
    from recbole_cdr_b200.data import Interaction
    from recbole_cdr_b200.trainer import GraphedTrainStep
    from recbole_cdr_b200.utils import get_model
    from fake_data import make_batch
    ds = FakeDataset(201, 300, 280, 1, 500, 450)
    cfg = base_config(embedding_size=64, reg_weight=0.01, mlp_hidden_size=[64, 32, 16, 8])
    torch.manual_seed(1)
    model = get_model('CoNet')(cfg, ds).to('cuda')
    assert model.stack_passes and model.dense_engine == 1

    def batch(seed):
        r = np.random.RandomState(seed)
        b = make_batch(ds, 'source', 1024, r)
        b.update(make_batch(ds, 'target', 1024, r))
        return Interaction(b).to('cuda')

    import copy
    eager = copy.deepcopy(model)
    prev = ops.set_cross_streams(streams)
    try:
        step = GraphedTrainStep(model, batch(0))
        for seed in (1, 2, 3):
            b = batch(seed)
            step.zero_table_grads()
            loss_g = step(b).clone()
            ops.set_cross_streams(0)
            eager.zero_grad(set_to_none=True)
            loss_e = eager.calculate_loss(b)
            loss_e.sum().backward()
            ops.set_cross_streams(streams)
            torch.cuda.synchronize()
            torch.testing.assert_close(loss_g, loss_e.detach().sum(), rtol=1e-6, atol=0)
            ge = dict(eager.named_parameters())
            for n, p in model.named_parameters():
                scale = max(1e-6, float(ge[n].grad.abs().max()))
                torch.testing.assert_close(p.grad, ge[n].grad, rtol=1e-4, atol=2e-5 * scale, msg=lambda s: f'{n}: {s}')
    finally:
        ops.set_cross_streams(prev)


def test_graphed_step_construction_leaves_model_and_optimizer_untouched():
    """The warm-up steps of GraphedTrainStep are real steps on the example batch; whatever they changed -- parameters, table
    gradients, optimizer state -- is put back, so the first replay is the first step of training."""
    from recbole_cdr_b200.data import Interaction
    from recbole_cdr_b200.trainer import GraphedTrainStep
    from recbole_cdr_b200.utils import get_model
    from fake_data import make_batch
    ds = FakeDataset(201, 300, 280, 101, 500, 450)
    cfg = base_config(embedding_size=64, mlp_hidden_size=[32, 16], dropout_prob=0.0, base_model='NeuMF', alpha=0.4)
    torch.manual_seed(1)
    model = get_model('DTCDR')(cfg, ds).to('cuda')
    import copy
    twin = copy.deepcopy(model)
    r = np.random.RandomState(3)
    b = make_batch(ds, 'source', 512, r)
    b.update(make_batch(ds, 'target', 512, r))
    b = Interaction(b).to('cuda')
    opt = torch.optim.SGD(model.parameters(), lr=1e-2, momentum=0.9)   # (graph-safe: its state is one tensor per parameter)
    step = GraphedTrainStep(model, b, optimizer=opt, warmup=3)
    for (n, p), (_, q) in zip(model.named_parameters(), twin.named_parameters()):
        assert torch.equal(p, q), n                                # the warm-up's optimizer steps were undone
    for st in opt.state.values():
        for k, v in st.items():
            if torch.is_tensor(v):
                assert float(v.abs().max()) == 0.0, k              # allocated (the graph points at it) and zeroed
    for p, g in step._table_grads.items():
        assert float(g.abs().max()) == 0.0
    # and the first replay equals the first eager step of the untouched twin
    opt2 = torch.optim.SGD(twin.parameters(), lr=1e-2, momentum=0.9)
    loss_g = step(b).clone()
    loss_e = twin.calculate_loss(b)
    loss_e.sum().backward()
    opt2.step()
    torch.testing.assert_close(loss_g, loss_e.detach().sum(), rtol=1e-6, atol=0)
    for (n, p), (_, q) in zip(model.named_parameters(), twin.named_parameters()):
        torch.testing.assert_close(p, q, rtol=1e-4, atol=1e-6, msg=lambda s: f'{n}: {s}')


def test_cmf_fused_both_epoch_tracks_per_batch_sgd():
    """CMF (the reference's default model): two weighted domain terms on shared tables through the persistent path."""
    from recbole_cdr_b200.utils import get_model, get_trainer, ModelType
    out = {}
    for fused in (0, 4):
        ds, loader = make_world(pairwise=False, seed=5)
        cfg = base_config(embedding_size=64, alpha=0.3, gamma=0.0, learner='sgd', learning_rate=20.0, weight_decay=0.0,
                          train_modes=['BOTH'], epoch_num=['3'], source_split=False, xdr_fused_steps=fused, **{'lambda': 0.0})
        torch.manual_seed(2022)
        model = get_model('CMF')(cfg, ds).to('cuda')
        trainer = get_trainer(ModelType.CROSSDOMAIN, 'CMF')(cfg, model)
        seen = []
        trainer.fit(loader, callback_fn=lambda epoch, loss: seen.append(loss))
        out[fused] = seen
    assert out[4][-1] < out[4][0] - 1e-3
    np.testing.assert_allclose(out[4], out[0], rtol=5e-3)


def test_device_pipeline_epoch_trains_without_host_batches():
    """Positives permuted and negatives drawn on the GPU -> [K, 3, B] blocks -> persistent launches with fused SGD."""
    from recbole_cdr_b200.data import DeviceDomainData
    from recbole_cdr_b200.sampler import CrossDomainSourceSampler
    from recbole_cdr_b200.utils import get_model, get_trainer, ModelType
    ds = FakeDataset(201, 300, 280, 1, 500, 450)
    rng = np.random.RandomState(0)
    su, si = ds.valid_ids('source')
    s_u, s_i = rng.choice(su, 20000), rng.choice(si, 20000)
    smp = CrossDomainSourceSampler('train', ds, user_ids=s_u, item_ids=s_i, device='cuda').set_phase('train')
    data = DeviceDomainData(s_u, s_i, smp)
    blocks = list(data.epoch_blocks(1024, 8, pairwise=True, generator=torch.Generator(device='cuda').manual_seed(1)))
    assert sum(b[0].shape[0] for b in blocks) == 20000 // 1024 and blocks[0][0].shape == (8, 3, 1024)
    used = set(zip(s_u.tolist(), s_i.tolist()))
    ids = blocks[0][0].cpu()
    valid_items = set(si.tolist())
    for u, n in zip(ids[:, 0].reshape(-1).tolist(), ids[:, 2].reshape(-1).tolist()):
        assert n in valid_items and (u, n) not in used          # the sampler's contract holds inside the pipeline
    pids, plab = next(iter(data.epoch_blocks(1024, 4, pairwise=False)))
    assert pids.shape == (4, 2, 1024) and plab.shape == (4, 1024) and plab[:, :512].all() and not plab[:, 512:].any()
    cfg = emcdr_cfg(learner='sgd', learning_rate=20.0, reg_weight=0.0, train_modes=['SOURCE'], epoch_num=['1'])
    torch.manual_seed(2022)
    model = get_model('EMCDR')(cfg, ds).to('cuda')
    model.set_phase('SOURCE')
    trainer = get_trainer(ModelType.CROSSDOMAIN, 'EMCDR')(cfg, model)
    losses = [trainer.train_epoch_device(data, 1024, steps_per_launch=8) for _ in range(4)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0] - 1e-3


def test_fused_runner_returns_one_loss_tensor_per_chunk():
    """with more chunks than device buffers in flight, the losses of chunk n must not be overwritten by chunk
    n + n_buffers (every run returns its own pinned tensor), and a change of the block shape (short last chunk) must not let
    the copy stream overwrite ids an in-flight launch still reads.  Checked against the same batches one launch at a time."""
    from recbole_cdr_b200 import ops
    from recbole_cdr_b200.trainer import FusedStepRunner
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(11)
    nu, ni, D, B, K = 5000, 7000, 64, 1024, 4
    ut, it = (torch.randn(nu, D, generator=g) * 0.1).to(dev), (torch.randn(ni, D, generator=g) * 0.1).to(dev)
    n_chunks = 7   # > 3 * n_buffers
    blocks = [torch.stack([torch.randint(1, nu, (K, B), generator=g), torch.randint(1, ni, (K, B), generator=g),
                           torch.randint(1, ni, (K, B), generator=g)], 1).pin_memory() for _ in range(n_chunks)]
    blocks.append(blocks[0][:2].clone().pin_memory())   # short last chunk: the buffers are re-created
    spec = dict(user_tab=ut, item_tab=it, pairwise=True, reg_weight=0.01, gamma=1e-10)
    gu, gi = torch.zeros_like(ut), torch.zeros_like(it)
    runner = FusedStepRunner(spec, lr=None, grad_tables=(gu, gi), n_buffers=2)
    got = [runner.run(b) for b in blocks]
    runner.synchronize()
    assert len({t.data_ptr() for t in got}) == len(got)
    gu2, gi2 = torch.zeros_like(ut), torch.zeros_like(it)
    for b, l in zip(blocks, got):
        d = b.to(dev)
        o, _, _ = ops.train_steps(ut, it, d[:, 0], d[:, 1], d[:, 2], reg_weight=0.01, user_dst=gu2, item_dst=gi2)
        torch.cuda.synchronize()
        assert torch.equal(l, o[:, 0].cpu())
    torch.testing.assert_close(gu, gu2, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(gi, gi2, rtol=1e-5, atol=1e-7)
