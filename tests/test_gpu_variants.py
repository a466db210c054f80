"""Hardware parity of the drop-in classes on the SECOND set of reference goldens (tests/golden/v_*.npz, outputs of the
unmodified reference classes: other row widths, ragged batches, duplicated ids, deeper stacks, reg_weight 0, 1- and 3-layer
BiTGCF, ``full_sort_predict``).  Default (composed, hardware-validated) kernels only; the same table through the kernels that
have not met hardware yet is in tests/test_gpu_engines.py.  CPU twin through the emulator: tests/test_emu_variants.py."""
import pytest

import variants_util as V
from golden_util import Golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', V.VARIANTS)
def test_variant_matches_the_reference(name):
    g = Golden(name)
    m = V.build(g, 'cuda')
    V.check_against_reference(m, g, 'cuda')


def test_conet_block_topk_matches_the_reference_scores():
    """CoNet.full_sort_topk (user blocks of the split-layer-0 tower + mask + topk; composed, hardware-validated kernels)
    against torch.topk of the reference's masked full_sort_predict scores."""
    g = Golden('v_conet_yaml_stack')
    m = V.build(g, 'cuda')
    V.check_topk_against_reference(m, g, 'cuda')
    m.full_sort_block_bytes = 4 * 64 * g.t('full_sort_predict').shape[1] * 2     # two users per block
    V.check_topk_against_reference(m, g, 'cuda')
