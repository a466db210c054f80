"""CPU tests of oracle/sampler_oracle.py: Philox known-answer vectors (Random123), the A0 layout, and the contract of
the negative draw (sampler/crossdomain_sampler.py:139-176, 212-213): uniform over candidates, never a used item,
`num` blocks of len(key_ids)."""
import numpy as np

from oracle import sampler_oracle as S


def test_philox4x32_10_known_answers():
    u = np.uint32
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = S.philox4x32_10(*(u(c) for c in ctr), *(u(k) for k in key))
        assert tuple(int(x) for x in got) == want


def test_joint_layout_matches_reference_ranges():
    lay = S.joint_layout(41, 30, 35)            # dataset.py:384-396 with 40 overlapped + PAD
    assert lay['overlap'] == (1, 41) and lay['target_only'] == (41, 71) and lay['source_only'] == (71, 106)
    assert lay['target_num'] == 71 and lay['source_num'] == 76 and lay['total'] == 106
    src = S.valid_ids(41, 30, 35, 'source')
    assert src[0] == 1 and 41 not in src and 70 not in src and src[-1] == 105 and len(src) == 75
    # candidate arithmetic == the reference's explicit list (crossdomain_sampler.py:212-213)
    ref_list = np.array(list(range(1, 41)) + list(range(41 + 30, 106)))
    assert np.array_equal(S.candidate_to_id(np.arange(len(ref_list)), 41, 30), ref_list)
    assert np.array_equal(S.valid_ids(41, 30, 35, 'target'), np.arange(1, 71))


def make_case(seed=0, n_users=50, n_ov=11, n_tgt=20, n_src=30, inter_per_user=12):
    rng = np.random.RandomState(seed)
    cand = S.valid_ids(n_ov, n_tgt, n_src, 'source')
    users = np.repeat(np.arange(1, n_users), inter_per_user)
    items = rng.choice(cand, size=len(users))
    rowptr, col = S.build_used_csr(users, items, n_users)
    return cand, rowptr, col, n_ov, n_tgt, len(cand)


def test_negative_draw_contract():
    cand, rowptr, col, n_ov, n_gap, n_valid = make_case()
    keys = np.random.RandomState(1).randint(1, 50, size=300)
    out, exhausted = S.neg_sample_uniform(keys, 3, rowptr, col, n_ov, n_gap, n_valid, seed=2022, stream_id=1)
    assert not exhausted and out.shape == (900,)
    assert np.isin(out, cand).all()                                  # only valid source-domain items
    tiled = np.tile(keys, 3)                                         # `num` blocks of len(keys)
    for u, v in zip(tiled, out):
        assert v not in col[rowptr[u]:rowptr[u + 1]]                 # never a used item
    again, _ = S.neg_sample_uniform(keys, 3, rowptr, col, n_ov, n_gap, n_valid, seed=2022, stream_id=1)
    assert np.array_equal(out, again)                                # counter-based: reproducible
    other, _ = S.neg_sample_uniform(keys, 3, rowptr, col, n_ov, n_gap, n_valid, seed=2022, stream_id=2)
    assert not np.array_equal(out, other)                            # a new call draws a new stream


def test_negative_draw_is_uniform_over_candidates():
    n_ov, n_gap, n_valid = 6, 4, 15                                  # ids 1..5 and 10..19
    rowptr, col = np.zeros(3, dtype=np.int64), np.zeros(0, dtype=np.int64)
    out, _ = S.neg_sample_uniform(np.ones(30000, dtype=np.int64), 1, rowptr, col, n_ov, n_gap, n_valid, 7, 0)
    cand = S.candidate_to_id(np.arange(n_valid), n_ov, n_gap)
    counts = np.array([(out == c).sum() for c in cand])
    assert counts.sum() == 30000
    chi2 = ((counts - 2000.0) ** 2 / 2000.0).sum()
    assert chi2 < 40.0                                               # 14 dof: p ~ 2e-4 at 40


def test_exhaustion_is_reported_not_looped_forever():
    n_ov, n_gap, n_valid = 4, 0, 3                                   # ids 1..3, user 1 used all of them
    rowptr, col = S.build_used_csr([1, 1, 1], [1, 2, 3], 2)
    out, exhausted = S.neg_sample_uniform([1], 1, rowptr, col, n_ov, n_gap, n_valid, 1, 0, max_attempts=20)
    assert exhausted
