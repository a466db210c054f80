"""The alternative engines of the tensor-core tile primitives (tc_tile.cuh, -DXDR_TC_MODE) under the CPU emulator:
mode 1 = bf16x3 on mma.m16n8k16 (parity-grade: same tolerances as the default 3xTF32 engine), mode 2 = one TF32 pass
(a diagnostic build: only checked to be a TF32-accurate version of the same computation)."""
import numpy as np
import pytest

import emu_util
import test_emu_conet as TC
import test_emu_mlp as TM


@pytest.mark.parametrize('batch,tile_rows', [(33, 32), (100, 64)])
def test_bf16x3_map_step_meets_the_parity_bar(batch, tile_rows):
    with emu_util.tc_mode(1):
        TM.test_map_step_matches_oracle(1, batch, tile_rows)


def test_bf16x3_dtcdr_term_meets_the_parity_bar():
    with emu_util.tc_mode(1):
        TM.test_dtcdr_term_matches_oracle(1, 129, 64)


@pytest.mark.parametrize('batch,dim,hidden,want,overlap_users', [(70, 32, [32, 16, 8], 0, True), (64, 128, [64, 32, 16, 8], 1, True),
                                                                 (33, 32, [16], 0, True)])
def test_bf16x3_conet_pass_meets_the_parity_bar(batch, dim, hidden, want, overlap_users):
    with emu_util.tc_mode(1):
        TC.test_conet_pass_matches_oracle(batch, dim, hidden, want, overlap_users)


def test_single_tf32_pass_is_the_same_computation_at_tf32_accuracy():
    src, tgt, ws, bs, idx, leaves, ref = TM.map_case(100)
    with emu_util.tc_mode(2):
        emu_util.config(sms=3, seed=0)
        r = emu_util.mlp_step(1, [64, 128, 64], [w.numpy() for w in ws], [b.numpy() for b in bs], TM.ACT_TANH, 0, 0,
                              (src.numpy(), None, None, None, tgt.numpy()), idx.numpy(), None, None, grad_loss=1.3, tile_rows=64)
    assert abs(r['loss'] - ref.item()) <= 5e-3 * abs(ref.item())
    g = leaves[2].grad.numpy()
    assert np.abs(r['dW'][0] - g).max() <= 2e-2 * np.abs(g).max()
    assert np.abs(r['dW'][0] - g).max() > 1e-7 * np.abs(g).max()      # and it really is the low-precision engine
