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
            pytest.skip('model scores with an MLP / propagated tables: no fused top-k entry')
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
