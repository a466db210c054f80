"""The drop-in classes on the SECOND set of reference goldens (tests/golden/v_*.npz: other row widths, ragged batch sizes,
heavily duplicated ids, deeper stacks, reg_weight 0, 1- and 3-layer BiTGCF, full_sort_predict), on CPU through the CTA emulator
(real kernel sources, tests/emu).  The same table runs on hardware in tests/test_gpu_variants.py."""
import pytest

import emu_util
import variants_util as V
from golden_util import Golden


def test_the_variant_table_is_complete():
    assert len(V.VARIANTS) == V.EXPECTED, V.VARIANTS


@pytest.mark.parametrize('name', V.VARIANTS)
def test_variant_matches_the_reference(name):
    g = Golden(name)
    with emu_util.patched_ops():
        m = V.build(g, 'cpu')
        V.check_against_reference(m, g, 'cpu')


@pytest.mark.parametrize('name', [n for n in V.VARIANTS if Golden(n).has('full_sort_predict')])
def test_variant_fused_topk_matches_the_reference_scores(name):
    g = Golden(name)
    with emu_util.patched_ops():
        m = V.build(g, 'cpu')
        if not hasattr(m, 'full_sort_topk'):
            pytest.skip('model scores with propagated tables: no top-k entry')
        V.check_topk_against_reference(m, g, 'cpu')
        if hasattr(m, 'full_sort_block_bytes'):     # CoNet: user blocks + mask + topk; also with two users per block
            m.full_sort_block_bytes = 4 * 64 * g.t('full_sort_predict').shape[1] * 2
            V.check_topk_against_reference(m, g, 'cpu')


@pytest.mark.parametrize('engine', ['fma', 'tc'])
@pytest.mark.parametrize('name', [n for n in V.VARIANTS if V.spec(Golden(n))['model'] in ('EMCDR', 'DTCDR')])
def test_variant_fused_mlp_engines(name, engine):
    """The fused row-tile kernels (fp32 FMA / tensor core) on the variants whose stack they support; an unsupported stack
    must fall back to the composed kernels and still match."""
    g = Golden(name)
    with emu_util.patched_ops():
        m = V.build(g, 'cpu', xdr_fused_mlp=engine)
        V.check_against_reference(m, g, 'cpu')


@pytest.mark.parametrize('name', [n for n in V.VARIANTS if V.spec(Golden(n))['model'] == 'CoNet'])
def test_variant_fused_conet(name):
    g = Golden(name)
    with emu_util.patched_ops():
        m = V.build(g, 'cpu', xdr_fused_conet=True)
        V.check_against_reference(m, g, 'cpu')


def test_conet_full_sort_predict_blocks_and_pairwise_predict_agree():
    """CoNet.full_sort_predict splits layer 0 into a per-item and a per-user half: any block size gives the reference's scores,
    and each entry equals ``predict`` on that (user, item) pair."""
    import torch
    g = Golden('v_conet_yaml_stack')
    with emu_util.patched_ops():
        m = V.build(g, 'cpu')
        fb = V.batch(g, 'cpu', 'fbatch/')
        ref = g.t('full_sort_predict')
        n_items = ref.shape[1]
        for block_bytes in (256 << 20, 4 * 64 * n_items * 3, 1):     # all users at once / three users per block / one by one
            m.full_sort_block_bytes = block_bytes
            torch.testing.assert_close(m.full_sort_predict(fb), ref, rtol=1e-4, atol=2e-6)
        users = fb['target_user_id']
        from recbole_cdr_b200.data import Interaction
        pairs = Interaction({'target_user_id': users.repeat_interleave(n_items),
                             'target_item_id': torch.arange(n_items).repeat(users.numel())})
        torch.testing.assert_close(m.full_sort_predict(fb).reshape(-1), m.predict(pairs).reshape(-1), rtol=1e-5, atol=1e-6)
