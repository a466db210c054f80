"""Hardware parity of the drop-in classes on the SECOND set of reference goldens (tests/golden/v_*.npz, outputs of the
unmodified reference classes: other row widths, ragged batches, duplicated ids, deeper stacks, reg_weight 0, 1- and 3-layer
BiTGCF, ``full_sort_predict``).  Default (composed, hardware-validated) kernels only; the same table through the kernels that
have not met hardware yet is in tests/test_gpu_unvalidated.py.  CPU twin through the emulator: tests/test_emu_variants.py."""
import pytest

import variants_util as V
from golden_util import Golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', V.VARIANTS)
def test_variant_matches_the_reference(name):
    g = Golden(name)
    m = V.build(g, 'cuda')
    V.check_against_reference(m, g, 'cuda')
