"""Graph operators of BiTGCF over libxdr: the normalised adjacency (A9), propagate (A10) and transfer+normalise
(A11-A12) with their backward passes."""
import numpy as np
import torch

from ._lib import call, cur_stream, ptr


def norm_adj_coo(rows, cols, n_users, n_items):
    """(row, col, value) of ``L = D^-1/2 A D^-1/2`` in CSR order (by row, then column) -- see ``NormAdj``."""
    n = n_users + n_items
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    keys = np.unique(rows * np.int64(n_items) + cols)          # de-duplicated (user, item) pairs, sorted by (user, item)
    eu, ei = keys // n_items, keys % n_items
    r = np.concatenate([eu, ei + n_users])
    c = np.concatenate([ei + n_users, eu])
    deg = np.bincount(r, minlength=n).astype(np.float64) + 1e-7
    dinv = np.power(deg, -0.5)
    val = (dinv[r] * 1.0 * dinv[c]).astype(np.float32)
    order = np.argsort(r * np.int64(n) + c, kind='stable')  # CSR order: by row, then column
    return r[order], c[order], val[order]


class NormAdj(object):
    """``L = D^-1/2 A D^-1/2`` of a bipartite interaction graph as CSR on the device + its work-item cut.

    Restates ``BiTGCF.get_norm_adj_mat`` (reference bitgcf.py:92-116): ``A = [[0, R], [R^T, 0]]`` over ``n_users + n_items``
    nodes with unit entries (duplicate interactions collapse to one, as the reference's dict does),
    ``D_ii = rowcount(A > 0) + 1e-7``, values ``float32(d_r^-1/2 * 1 * d_c^-1/2)`` with the inverse roots in float64 (SciPy's
    ``D * A * D`` followed by ``torch.FloatTensor``).  One-off host work at model construction (the reference builds a
    Python dict of every nonzero and uses the private ``dok_matrix._update``, gone from SciPy >= 1.13).  L is symmetric,
    so the same CSR serves the backward pass.
    """

    def __init__(self, rows, cols, n_users, n_items, device, chunk=256):
        n = n_users + n_items
        r, c, val = norm_adj_coo(rows, cols, n_users, n_items)
        self.n, self.n_users, self.n_items = n, n_users, n_items
        self.device = torch.device(device)
        self._set_csr(r, c, val, n, chunk)

    def _set_csr(self, r, c, val, n_rows, chunk):
        """r (sorted), c, val: the nonzeros of the rows this object holds -> CSR + work-item cut on the device"""
        rowptr = np.zeros(n_rows + 1, dtype=np.int64)
        np.add.at(rowptr, r + 1, 1)
        rowptr = np.cumsum(rowptr)
        # work items: every row is cut into pieces of <= chunk nonzeros (rows without nonzeros still get one empty item
        # so that their output row is written)
        counts = rowptr[1:] - rowptr[:-1]
        pieces = np.maximum(1, -(-counts // chunk))
        work_row = np.repeat(np.arange(n_rows, dtype=np.int64), pieces)
        first = np.cumsum(pieces) - pieces
        k = np.arange(work_row.size, dtype=np.int64) - np.repeat(first, pieces)
        work_beg = rowptr[work_row] + k * chunk
        work_end = np.minimum(work_beg + chunk, rowptr[work_row + 1])
        work_split = (pieces[work_row] > 1).astype(np.uint8)
        split_rows = np.nonzero(pieces > 1)[0].astype(np.int64)
        self.n_rows, self.nnz = n_rows, int(val.size)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.rowptr, self.col, self.val = t(rowptr), t(c), t(val)
        self.work_row, self.work_beg, self.work_end, self.work_split = t(work_row), t(work_beg), t(work_end), t(work_split)
        self.split_rows = t(split_rows)

    def to_sparse_coo(self):
        """torch sparse COO view of L (tests)."""
        r = torch.repeat_interleave(torch.arange(self.n, device=self.device), self.rowptr[1:] - self.rowptr[:-1])
        return torch.sparse_coo_tensor(torch.stack([r, self.col]), self.val, (self.n, self.n)).coalesce()

    def spmm(self, X):
        """S = L . X  (torch.sparse.mm(L, X), bitgcf.py:131)"""
        X = X.contiguous()
        S = torch.empty_like(X)
        call('xdr_spmm_csr', ptr(self.work_row), ptr(self.work_beg), ptr(self.work_end), ptr(self.work_split),
             self.work_row.numel(), ptr(self.split_rows), self.split_rows.numel(), ptr(self.col), ptr(self.val), ptr(X),
             X.shape[1], ptr(S), cur_stream())
        return S


def _elementwise(A, B, C, mode):
    out = torch.empty_like(A)
    call('xdr_prop_elementwise', ptr(A), ptr(B), ptr(C), ptr(out), A.numel(), mode, cur_stream())
    return out


class GraphProp(torch.autograd.Function):
    """``BiTGCF.graph_layer`` with ``drop_rate = 0`` (bitgcf.py:130-135): ``S = L.E;  E' = E + S + E*S``.
    Backward (L symmetric): ``dE = dE'*(1 + S) + L.(dE'*(1 + E))``."""

    @staticmethod
    def forward(ctx, E, adj):
        E = E.contiguous()
        S = adj.spmm(E)
        ctx.adj = adj
        ctx.save_for_backward(E, S)
        return _elementwise(E, S, None, 0)

    @staticmethod
    def backward(ctx, G):
        E, S = ctx.saved_tensors
        G = G.contiguous()
        dS = _elementwise(G, E, None, 1)
        T = ctx.adj.spmm(dS)
        return _elementwise(G, S, T, 2), None


class TransferNorm(torch.autograd.Function):
    """``BiTGCF.transfer_layer`` (bitgcf.py:137-172) on both domains + ``F.normalize(p=2, dim=1)`` (bitgcf.py:185-186).
    Returns (Es, Et, Ns, Nt): the transferred tables (next layer's input) and their row-normalised copies."""

    @staticmethod
    def forward(ctx, Ps, Pt, deg_s, deg_t, n_users, n_items, n_ov_users, n_ov_items, lam_s, lam_t):
        Ps, Pt = Ps.contiguous(), Pt.contiguous()
        Es, Et, Ns, Nt = (torch.empty_like(Ps) for _ in range(4))
        d = Ps.shape[1]
        call('xdr_transfer_norm_fwd', ptr(Ps), ptr(Pt), n_users, n_items, n_ov_users, n_ov_items, d, float(lam_s),
             float(lam_t), ptr(deg_s), ptr(deg_t), ptr(Es), ptr(Et), ptr(Ns), ptr(Nt), d, cur_stream())
        ctx.save_for_backward(Es, Et, deg_s, deg_t)
        ctx.meta = (n_users, n_items, n_ov_users, n_ov_items, float(lam_s), float(lam_t))
        return Es, Et, Ns, Nt

    @staticmethod
    def backward(ctx, dEs, dEt, dNs, dNt):
        Es, Et, deg_s, deg_t = ctx.saved_tensors
        n_users, n_items, n_ov_users, n_ov_items, lam_s, lam_t = ctx.meta
        d = Es.shape[1]
        zeros = None
        def dense(g):
            nonlocal zeros
            if g is None:
                if zeros is None:
                    zeros = torch.zeros_like(Es)
                return zeros
            return g.contiguous()
        dNs, dNt = dense(dNs), dense(dNt)
        has_e = dEs is not None or dEt is not None
        dEs2, dEt2 = (dense(dEs), dense(dEt)) if has_e else (None, None)
        dPs, dPt = torch.empty_like(Es), torch.empty_like(Et)
        call('xdr_transfer_norm_bwd', ptr(Es), ptr(Et), ptr(dNs), ptr(dNt), d, ptr(dEs2), ptr(dEt2), n_users, n_items,
             n_ov_users, n_ov_items, d, lam_s, lam_t, ptr(deg_s), ptr(deg_t), ptr(dPs), ptr(dPt), cur_stream())
        return dPs, dPt, None, None, None, None, None, None, None, None
