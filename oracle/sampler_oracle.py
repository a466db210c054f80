"""CPU oracle for the integer side of the hot path: joint id layout (A0) and the negative draw (A18).
TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Reference behaviour restated (paths relative to /root/reference/recbole_cdr/):

* candidate list of the source-domain sampler, sampler/crossdomain_sampler.py:212-213:
      item_id_list = [1, n_ov_items) ++ [n_ov_items + n_tgt_only_items, n_total_items)
  target-domain draws use recbole's own Sampler: uniform over [1, item_num)  [recbole-1.0.1].
* sample_by_key_ids, sampler/crossdomain_sampler.py:139-176: draw uniformly from the candidates; every position whose
  draw is in used_ids[user] is redrawn until none is; output laid out as `num` blocks of len(key_ids)
  (key_ids = np.tile(key_ids, num), :166).
* used_ids (get_used_ids, :229-250): per user the set of items it interacted with.

The reference draws with NumPy's global MT19937, so bitwise RNG parity with it is not a goal;
the contract is: uniform over the candidates AND never in the user's used set AND the reference's output layout.
To make the CUDA kernel checkable bit for bit, both sides use the same counter-based generator, Philox4x32-10
(Salmon et al., SC'11): key = seed, counter = (position, attempt, stream).  Parity pinning: the generator is checked
against the published Random123 known-answer vectors; the sampler is checked by its properties.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs: uint32 arrays (broadcastable); returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.asarray(k0, dtype=np.uint32)
    k1 = np.asarray(k1, dtype=np.uint32)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return c0, c1, c2, c3


def _mulhi64(a, b):
    """High 64 bits of the 128-bit product of uint64 arrays a and scalar b (python ints keep it exact)."""
    return np.array([(int(x) * int(b)) >> 64 for x in a], dtype=np.int64)


def candidate_to_id(k, n_overlap, n_gap):
    """k-th valid id: k+1 if k+1 < n_overlap else k+1+n_gap  (n_gap = n_tgt_only for the source domain, 0 for the target)."""
    k1 = np.asarray(k, dtype=np.int64) + 1
    return np.where(k1 < n_overlap, k1, k1 + n_gap)


def draw(pos, attempt, n_valid, n_overlap, n_gap, seed, stream_id):
    """One uniform draw per (position, attempt): 64 random bits -> floor(r * n_valid / 2^64) -> joint id."""
    pos = np.asarray(pos, dtype=np.uint64)
    r0, r1, _, _ = philox4x32_10((pos & MASK32).astype(np.uint32), (pos >> np.uint64(32)).astype(np.uint32),
                                 np.asarray(attempt, dtype=np.uint32), np.uint32(stream_id & 0xFFFFFFFF),
                                 np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
    r = r0.astype(np.uint64) | (r1.astype(np.uint64) << np.uint64(32))
    return candidate_to_id(_mulhi64(r, n_valid), n_overlap, n_gap)


def neg_sample_uniform(key_ids, num, used_rowptr, used_col, n_overlap, n_gap, n_valid, seed, stream_id, max_attempts=1000):
    """sample_by_key_ids (crossdomain_sampler.py:139-176) with the Philox stream above.
    Returns (value_ids [len(key_ids)*num] int64 laid out as `num` blocks of len(key_ids), exhausted flag)."""
    key_ids = np.asarray(key_ids, dtype=np.int64)
    P = len(key_ids)
    total = P * num
    keys = np.tile(key_ids, num)                      # crossdomain_sampler.py:166
    out = np.zeros(total, dtype=np.int64)
    check = np.arange(total)
    attempt = np.zeros(total, dtype=np.int64)
    exhausted = False
    while len(check) > 0:
        out[check] = draw(check, attempt[check], n_valid, n_overlap, n_gap, seed, stream_id)
        still = []
        for i in check:
            u = keys[i]
            row = used_col[used_rowptr[u]:used_rowptr[u + 1]]
            j = np.searchsorted(row, out[i])
            if j < len(row) and row[j] == out[i]:
                attempt[i] += 1
                if attempt[i] >= max_attempts:
                    exhausted = True
                else:
                    still.append(i)
        check = np.array(still, dtype=np.int64)
    return out, exhausted


def build_used_csr(user_ids, item_ids, n_users):
    """get_used_ids (crossdomain_sampler.py:229-250) as CSR: per user the sorted, de-duplicated items it interacted with."""
    user_ids = np.asarray(user_ids, dtype=np.int64)
    item_ids = np.asarray(item_ids, dtype=np.int64)
    pairs = np.unique(np.stack([user_ids, item_ids], 1), axis=0) if len(user_ids) else np.zeros((0, 2), np.int64)
    rowptr = np.zeros(n_users + 1, dtype=np.int64)
    np.add.at(rowptr, pairs[:, 0] + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, pairs[:, 1].copy()


# ---- A0: joint id layout (data/dataset.py:344-445) -------------------------------------------------------------------

def joint_layout(n_overlap, n_target_only, n_source_only):
    """Ranges of the joint id space: 0 = PAD, [1, n_ov) overlapped, then target-only, then source-only."""
    t0 = n_overlap
    s0 = n_overlap + n_target_only
    total = s0 + n_source_only
    return {'pad': 0, 'overlap': (1, n_overlap), 'target_only': (t0, s0), 'source_only': (s0, total), 'total': total,
            'target_num': n_overlap + n_target_only, 'source_num': n_overlap + n_source_only}


def valid_ids(n_overlap, n_target_only, n_source_only, domain):
    lay = joint_layout(n_overlap, n_target_only, n_source_only)
    if domain == 'source':
        return np.concatenate([np.arange(*lay['overlap']), np.arange(*lay['source_only'])])
    return np.arange(1, lay['target_num'])
