"""TEST INFRASTRUCTURE (oracle): numpy restatement of full-sort evaluation scoring for one block of users.

Follows the reference's ``full_sort_predict`` (``emcdr.py:208-233`` / ``cmf.py:107-112``: ``matmul(user_e, all_item_e.T)``
over the target items) and the masking recbole's full-sort evaluation applies before ``torch.topk`` [recbole-1.0.1
``FullSortEvalDataLoader`` history index + ``Collector``: PAD column and history set to -inf].  Tie order, which torch
leaves unspecified, is fixed to ascending item id.  Only ``tests/`` may import this module."""
import numpy as np


def full_sort_scores(user_vecs, item_tab, n_items=None):
    """[B, n_items] fp32 scores, the reference's dense matrix."""
    n_items = item_tab.shape[0] if n_items is None else n_items
    return (user_vecs.astype(np.float32) @ item_tab[:n_items].astype(np.float32).T).astype(np.float32)


def masked_topk(scores, k, first_item=1, hist_ptr=None, hist_ids=None):
    """Top-k per row after masking items < first_item and each row's history; returns (scores [B, k], ids [B, k]) with
    (-inf, -1) padding, ordered by score descending then id ascending."""
    s = scores.astype(np.float32).copy()
    s[:, :first_item] = -np.inf
    B, N = s.shape
    if hist_ptr is not None:
        for u in range(B):
            s[u, hist_ids[hist_ptr[u]:hist_ptr[u + 1]]] = -np.inf
    out_s = np.full((B, k), -np.inf, dtype=np.float32)
    out_i = np.full((B, k), -1, dtype=np.int64)
    ids = np.arange(N)
    for u in range(B):
        order = np.lexsort((ids, -s[u]))          # primary: score descending; secondary: id ascending
        order = order[np.isfinite(s[u][order])][:k]
        out_s[u, :len(order)] = s[u][order]
        out_i[u, :len(order)] = order
    return out_s, out_i
