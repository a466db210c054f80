"""CPU tests of the batch multiplexer (A19) against the behaviour of reference data/dataloader.py:99-180."""
import numpy as np
import pytest
import torch

from recbole_cdr_b200.data import CrossDomainDataloader, DomainTrainDataLoader, Interaction, OverlapDataloader
from recbole_cdr_b200.utils import CrossDomainDataLoaderState as S


class StubSampler:
    """Deterministic stand-in: 'negative' of item i is i + 1000 (the multiplexer must not care)."""

    def sample_by_user_ids(self, user_ids, item_ids, num):
        return torch.as_tensor(item_ids).repeat(num) + 1000


def loaders(n_src=10, n_tgt=7, bs=4, pairwise=False):
    src = DomainTrainDataLoader('source_user_id', 'source_item_id', np.arange(1, n_src + 1), np.arange(101, 101 + n_src), bs,
                                StubSampler(), pairwise, 'source_label')
    tgt = DomainTrainDataLoader('target_user_id', 'target_item_id', np.arange(1, n_tgt + 1), np.arange(201, 201 + n_tgt), bs,
                                StubSampler(), pairwise, 'target_label')
    ov = OverlapDataloader(9, 4, generator=torch.Generator().manual_seed(0))
    return CrossDomainDataloader(src, tgt, ov)


def test_pointwise_expansion_layout():
    dl = loaders()
    dl.set_mode(S.SOURCE)
    b = next(iter(dl))
    assert isinstance(b, Interaction) and set(b.columns) == {'source_user_id', 'source_item_id', 'source_label'}
    # step = batch_size // 2 positives, then the same users with sampled items, labels 1..1,0..0 (fp32), ids int64
    assert b['source_user_id'].tolist() == [1, 2, 1, 2] and b['source_item_id'].tolist() == [101, 102, 1101, 1102]
    assert b['source_label'].tolist() == [1.0, 1.0, 0.0, 0.0] and b['source_label'].dtype == torch.float32
    assert b['source_user_id'].dtype == torch.int64 and len(dl) == 5


def test_pairwise_expansion_layout():
    dl = loaders(pairwise=True)
    dl.set_mode(S.TARGET)
    b = next(iter(dl))
    assert set(b.columns) == {'target_user_id', 'target_item_id', 'neg_target_item_id'}
    assert b['target_user_id'].tolist() == [1, 2, 3, 4] and b['neg_target_item_id'].tolist() == [1201, 1202, 1203, 1204]


def test_both_state_epoch_follows_target_and_source_wraps():
    dl = loaders(n_src=3, n_tgt=7, bs=4)           # source: 2 batches per pass, target: 4 batches per epoch
    dl.set_mode(S.BOTH)
    assert len(dl) == 4 and dl.pr_end == 7
    batches = list(dl)
    assert len(batches) == 4                        # the epoch ends with the TARGET loader (dataloader.py:119-123)
    for b in batches:
        assert {'source_user_id', 'source_item_id', 'source_label', 'target_user_id', 'target_item_id', 'target_label'} \
            == set(b.columns)
    # the source loader silently restarted (dataloader.py:155-159): batch 2 is its first batch again
    assert batches[2]['source_user_id'].tolist() == batches[0]['source_user_id'].tolist()
    # ragged halves on the last batch: target has 1 positive left, source has a full step
    assert batches[3]['target_user_id'].numel() == 2 and batches[3]['source_user_id'].numel() in (2, 4)
    assert dl.source_dataloader.pr == 0 and dl.target_dataloader.pr == 0      # reset for the next epoch
    assert len(list(dl)) == 4


def test_overlap_state_yields_column_vectors_including_pad():
    dl = loaders()
    dl.set_mode(S.OVERLAP)
    bs = list(dl)
    assert [tuple(b['overlap'].shape) for b in bs] == [(4, 1), (4, 1), (1, 1)]
    assert sorted(torch.cat([b['overlap'] for b in bs]).reshape(-1).tolist()) == list(range(9))   # PAD id 0 included


def test_state_change_rules():
    dl = loaders()
    with pytest.raises(NotImplementedError):
        dl.set_mode('BOTH')                          # not a CrossDomainDataLoaderState
    dl.set_mode(S.SOURCE)
    it = iter(dl)
    next(it)
    with pytest.raises(PermissionError):
        dl.set_mode(S.TARGET)                        # mid-epoch (dataloader.py:177-179)
    for _ in it:
        pass
    dl.set_mode(S.TARGET)


def test_interaction_update_and_to():
    a = Interaction({'x': torch.arange(3)})
    b = Interaction({'y': torch.arange(5).float()})
    a.update(b)
    assert len(a) == 5 and set(a.columns) == {'x', 'y'} and a.to('cpu')['y'].dtype == torch.float32
