"""Autograd-visible operators over libxdr (PyTorch tensors in, PyTorch tensors out).

Each ``torch.autograd.Function`` here is the fused replacement of a short chain of ATen ops on the reference's hot
path; the docstring of each names the reference lines.  Forward and backward both run hand-written CUDA through the
C ABI (``_lib.call``); PyTorch only allocates the output tensors and supplies the stream.

Table gradients come in two modes (``set_table_grad_mode``):

* ``'autograd'`` (default, exact drop-in): backward returns a dense ``[N, D]`` gradient per table -- zero-filled, then
  scatter-added by the kernel -- exactly what ``embedding_dense_backward`` produces in the reference, so any
  ``torch.optim`` optimizer and ``torch.autograd.grad`` work unchanged.
* ``'inplace'``: backward scatter-adds straight into ``table.grad`` (allocated on first use) and returns ``None`` for
  the table, skipping one dense zero-fill and one dense add per use of a table.  ``loss.backward()`` leaves the same
  ``.grad`` contents; ``torch.autograd.grad`` does not see table gradients in this mode.
"""
import ctypes as _ct
import os as _os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import call, cur_stream, ptr

_TABLE_GRAD_MODE = 'autograd'
# tcgen05 map step + 'inplace' table gradients: loss and gradients in ONE launch at forward time (XDR_EAGER_TC5=0 switches it off)
EAGER_TC5 = _os.environ.get('XDR_EAGER_TC5', '1') != '0'
# set True (or env XDR_CHECK_IDS=1) to raise IndexError on out-of-range ids (adds a sync per op)
CHECK_IDS = _os.environ.get('XDR_CHECK_IDS', '0') == '1'


def set_table_grad_mode(mode: str):
    global _TABLE_GRAD_MODE
    if mode not in ('autograd', 'inplace'):
        raise ValueError("table grad mode must be 'autograd' or 'inplace'")
    _TABLE_GRAD_MODE = mode


def get_table_grad_mode() -> str:
    return _TABLE_GRAD_MODE


def _require_cuda_f32(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor: the xdr hot path has no CPU implementation')
    if t.dtype != torch.float32:
        raise TypeError(f'{name} must be float32, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name} must be contiguous')


def _on_device(t: torch.Tensor) -> bool:
    """True for tensors the kernels can address (CUDA memory)."""
    return t.is_cuda


def _ids(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor')
    if t.dtype != torch.int64:
        raise TypeError(f'{name} must be int64 (torch.LongTensor), got {t.dtype}')
    return t.contiguous()


def _grad_dst(table: torch.Tensor):
    """Destination for a table gradient in the current mode -> (dst tensor, value to return from backward)."""
    if _TABLE_GRAD_MODE == 'inplace' and table.is_leaf:
        if table.grad is None:
            table.grad = torch.zeros_like(table)
        return table.grad, None
    g = torch.zeros_like(table)
    return g, g


def _maybe_check(device):
    if CHECK_IDS:
        _lib.check_ids(device)


def _oob(device):
    """The device-side out-of-range flag is ALWAYS armed (a kernel that meets a bad id sets it; that costs nothing otherwise);
    ``CHECK_IDS`` only decides whether every op also reads it back (a sync per op).  ``check_ids_now`` reads it on demand --
    the trainer does so once per epoch, where it reads the loss anyway."""
    return ptr(_lib.oob_flag(device))


def check_ids_now(device):
    """Raise IndexError (as nn.Embedding would have, at the op) if any kernel since the last check met an out-of-range id."""
    _lib.check_ids(torch.device(device))


# ------------------------------------------------------------------------------------------------------------------
# A1: gather / scatter-add
# ------------------------------------------------------------------------------------------------------------------

def gather_rows_raw(table: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None, col: int = 0):
    """out[:, col:col+D] = table[idx] (bit-exact); ``out`` may be a wider [n, ld] buffer (the concat target)."""
    _require_cuda_f32(table, 'table')
    idx = _ids(idx, 'idx').reshape(-1)
    n, d = idx.numel(), table.shape[1]
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=table.device)
    ld = out.stride(0) if n > 0 else max(out.shape[1], d)
    view = out[:, col:col + d]
    call('xdr_gather_rows', ptr(table), table.shape[0], d, ptr(idx), n, view.data_ptr() if n else None, ld,
         _oob(table.device), cur_stream())
    _maybe_check(table.device)
    return out


def scatter_add_rows_raw(dst: torch.Tensor, idx: torch.Tensor, rows: torch.Tensor, scale: float = 1.0, col: int = 0):
    """dst[idx] += scale * rows[:, col:col+D]."""
    _require_cuda_f32(dst, 'dst')
    idx = _ids(idx, 'idx').reshape(-1)
    n, d = idx.numel(), dst.shape[1]
    if n == 0:
        return dst
    view = rows[:, col:col + d]
    call('xdr_scatter_add_rows', ptr(dst), dst.shape[0], d, ptr(idx), n, view.data_ptr(), rows.stride(0), float(scale),
         _oob(dst.device), cur_stream())
    return dst


class GatherRows(torch.autograd.Function):
    """``table[idx]`` == ``nn.Embedding.__call__`` (emcdr.py:99-100, conet.py:106-109, bitgcf.py:221-224);
    backward == ``embedding_dense_backward`` as a vector-atomic scatter-add."""

    @staticmethod
    def forward(ctx, table, idx):
        out = gather_rows_raw(table, idx)
        ctx.save_for_backward(table, idx)
        return out.view(*idx.shape, table.shape[1])

    @staticmethod
    def backward(ctx, grad_out):
        table, idx = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:   # a frozen table: no dense [N, D] gradient is allocated for it, nothing is scattered
            return None, None
        go = grad_out.reshape(-1, table.shape[1]).contiguous()
        dst, ret = _grad_dst(table)
        scatter_add_rows_raw(dst, idx, go)
        return ret, None


def gather_rows(table, idx):
    return GatherRows.apply(table, idx)


class GatherConcat(torch.autograd.Function):
    """``torch.cat([user_tab[user], item_tab[item]], dim=1)`` written in one pass per table straight into the
    concatenated buffer (conet.py:106-111)."""

    @staticmethod
    def forward(ctx, user_tab, item_tab, user, item):
        n, d = user.numel(), user_tab.shape[1]
        out = torch.empty((n, 2 * d), dtype=torch.float32, device=user_tab.device)
        gather_rows_raw(user_tab, user, out, 0)
        gather_rows_raw(item_tab, item, out, d)
        ctx.save_for_backward(user_tab, item_tab, user, item)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        user_tab, item_tab, user, item = ctx.saved_tensors
        go = grad_out.contiguous()
        d = user_tab.shape[1]
        ru = ri = None
        if ctx.needs_input_grad[0]:   # (frozen tables get neither a dense gradient nor a scatter)
            du, ru = _grad_dst(user_tab)
            scatter_add_rows_raw(du, user, go, 1.0, 0)
        if ctx.needs_input_grad[1]:
            di, ri = _grad_dst(item_tab)
            scatter_add_rows_raw(di, item, go, 1.0, d)
        return ru, ri, None, None


class GatherMax2Concat(torch.autograd.Function):
    """DTCDR.neumf_forward input: ``cat(max(Es_u[u], Et_u[u]), max(Es_i[i], Et_i[i]))`` (dtcdr.py:113-121)."""

    @staticmethod
    def forward(ctx, su_tab, tu_tab, si_tab, ti_tab, user, item):
        user, item = _ids(user, 'user'), _ids(item, 'item')
        n, d = user.numel(), su_tab.shape[1]
        for t, nm in ((su_tab, 'source_user'), (tu_tab, 'target_user'), (si_tab, 'source_item'), (ti_tab, 'target_item')):
            _require_cuda_f32(t, nm)
        out = torch.empty((n, 2 * d), dtype=torch.float32, device=su_tab.device)
        s = cur_stream()
        call('xdr_gather_max2', ptr(su_tab), ptr(tu_tab), su_tab.shape[0], d, ptr(user), n, out.data_ptr(), 2 * d,
             _oob(out.device), s)
        call('xdr_gather_max2', ptr(si_tab), ptr(ti_tab), si_tab.shape[0], d, ptr(item), n, out[:, d:].data_ptr(), 2 * d,
             _oob(out.device), s)
        _maybe_check(out.device)
        ctx.save_for_backward(su_tab, tu_tab, si_tab, ti_tab, user, item)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        su_tab, tu_tab, si_tab, ti_tab, user, item = ctx.saved_tensors
        go = grad_out.contiguous()
        n, d = user.numel(), su_tab.shape[1]
        s = cur_stream()
        dsu, rsu = _grad_dst(su_tab)
        dtu, rtu = _grad_dst(tu_tab)
        dsi, rsi = _grad_dst(si_tab)
        dti, rti = _grad_dst(ti_tab)
        call('xdr_scatter_max2_bwd', ptr(su_tab), ptr(tu_tab), su_tab.shape[0], d, ptr(user), n, go.data_ptr(), 2 * d, 1.0,
             ptr(dsu), ptr(dtu), None, s)
        call('xdr_scatter_max2_bwd', ptr(si_tab), ptr(ti_tab), si_tab.shape[0], d, ptr(item), n, go[:, d:].data_ptr(), 2 * d,
             1.0, ptr(dsi), ptr(dti), None, s)
        return rsu, rtu, rsi, rti, None, None


# ------------------------------------------------------------------------------------------------------------------
# A2/A3/A13/A16: fused gather -> score -> loss
# ------------------------------------------------------------------------------------------------------------------

class BprLoss(torch.autograd.Function):
    """Fused ``BPRLoss(s(u,i+), s(u,i-)) + reg_weight * EmbLoss(Eu[u], Ei[i+])`` -- EMCDR.calculate_source_loss /
    calculate_target_loss, BPR branch (emcdr.py:121-130, 144-153).  Returns the loss with shape ``[1]`` as the
    reference does."""

    @staticmethod
    def forward(ctx, user_tab, item_tab, user, pos_item, neg_item, gamma, reg_weight):
        _require_cuda_f32(user_tab, 'user table')
        _require_cuda_f32(item_tab, 'item table')
        user, pos_item, neg_item = _ids(user, 'user'), _ids(pos_item, 'pos_item'), _ids(neg_item, 'neg_item')
        b, d = user.numel(), user_tab.shape[1]
        dev = user_tab.device
        scores = torch.empty((2, b), dtype=torch.float32, device=dev)
        out8 = torch.empty(8, dtype=torch.float32, device=dev)
        call('xdr_bpr_fwd', ptr(user_tab), ptr(item_tab), user_tab.shape[0], item_tab.shape[0], d, ptr(user), ptr(pos_item),
             ptr(neg_item), b, float(gamma), float(reg_weight), scores[0].data_ptr(), scores[1].data_ptr(), ptr(out8),
             ptr(_lib.workspace(dev)), _oob(dev), cur_stream())
        _maybe_check(dev)
        ctx.save_for_backward(user_tab, item_tab, user, pos_item, neg_item, scores, out8)
        ctx.gamma, ctx.reg_weight = float(gamma), float(reg_weight)
        return out8[0:1].clone()

    @staticmethod
    def backward(ctx, grad_loss):
        user_tab, item_tab, user, pos_item, neg_item, scores, out8 = ctx.saved_tensors
        g = grad_loss.reshape(-1)[:1].contiguous().float()
        du, ru = _grad_dst(user_tab)
        di, ri = _grad_dst(item_tab)
        call('xdr_bpr_bwd', ptr(user_tab), ptr(item_tab), user_tab.shape[0], item_tab.shape[0], user_tab.shape[1], ptr(user),
             ptr(pos_item), ptr(neg_item), user.numel(), ctx.gamma, ctx.reg_weight, scores[0].data_ptr(),
             scores[1].data_ptr(), ptr(out8), ptr(g), 1.0, ptr(du), ptr(di), cur_stream())
        return ru, ri, None, None, None, None, None


def bpr_loss(user_tab, item_tab, user, pos_item, neg_item, reg_weight, gamma=1e-10):
    return BprLoss.apply(user_tab, item_tab, user, pos_item, neg_item, gamma, reg_weight)


class PointLoss(torch.autograd.Function):
    """Fused ``data_loss(dot(Eu[u], Ei[i]), label) + reg_weight * EmbLoss(Eu[u], Ei[i])`` with data_loss MSE
    (EMCDR-MF, emcdr.py:111-120), BCE(sigmoid(.)) (CMF cmf.py:75-98; BiTGCF bitgcf.py:226-228) or none (the
    EmbLoss-only term of bitgcf.py:231-233).  Returns ``[1]``."""

    @staticmethod
    def forward(ctx, user_tab, item_tab, user, item, label, loss_kind, reg_weight):
        _require_cuda_f32(user_tab, 'user table')
        _require_cuda_f32(item_tab, 'item table')
        user, item = _ids(user, 'user'), _ids(item, 'item')
        if label is not None:
            _require_cuda_f32(label, 'label')
        b, d = user.numel(), user_tab.shape[1]
        dev = user_tab.device
        score = torch.empty(b, dtype=torch.float32, device=dev)
        out8 = torch.empty(8, dtype=torch.float32, device=dev)
        call('xdr_point_fwd', ptr(user_tab), ptr(item_tab), user_tab.shape[0], item_tab.shape[0], d, ptr(user), ptr(item),
             ptr(label), b, int(loss_kind), float(reg_weight), ptr(score), ptr(out8), ptr(_lib.workspace(dev)), _oob(dev),
             cur_stream())
        _maybe_check(dev)
        ctx.save_for_backward(user_tab, item_tab, user, item, label, score, out8)
        ctx.loss_kind, ctx.reg_weight = int(loss_kind), float(reg_weight)
        return out8[0:1].clone()

    @staticmethod
    def backward(ctx, grad_loss):
        user_tab, item_tab, user, item, label, score, out8 = ctx.saved_tensors
        g = grad_loss.reshape(-1)[:1].contiguous().float()
        du, ru = _grad_dst(user_tab)
        di, ri = _grad_dst(item_tab)
        call('xdr_point_bwd', ptr(user_tab), ptr(item_tab), user_tab.shape[0], item_tab.shape[0], user_tab.shape[1],
             ptr(user), ptr(item), ptr(label), user.numel(), ctx.loss_kind, ctx.reg_weight, ptr(score), ptr(out8), ptr(g),
             1.0, ptr(du), ptr(di), cur_stream())
        return ru, ri, None, None, None, None, None


def point_loss(user_tab, item_tab, user, item, label, loss_kind, reg_weight):
    return PointLoss.apply(user_tab, item_tab, user, item, label, loss_kind, reg_weight)


def dot_score(user_tab, item_tab, user, item):
    """Inference-only ``(Eu[u] * Ei[i]).sum(1)`` (emcdr.py:98-108): the fused forward with no loss term."""
    with torch.no_grad():
        user, item = _ids(user, 'user'), _ids(item, 'item')
        b, dev = user.numel(), user_tab.device
        score = torch.empty(b, dtype=torch.float32, device=dev)
        out8 = torch.empty(8, dtype=torch.float32, device=dev)
        if b:
            call('xdr_point_fwd', ptr(user_tab), ptr(item_tab), user_tab.shape[0], item_tab.shape[0], user_tab.shape[1],
                 ptr(user), ptr(item), None, b, _lib.LOSS_NONE, 0.0, ptr(score), ptr(out8), ptr(_lib.workspace(dev)),
                 _oob(dev), cur_stream())
            _maybe_check(dev)
        return score


# ------------------------------------------------------------------------------------------------------------------
# A4/A7/A14: dense layers
# ------------------------------------------------------------------------------------------------------------------

class dense_engine(object):
    """``with ops.dense_engine(e):`` -- the dense entry points called inside run on engine ``e`` (``xdr_set_dense_engine``:
    0 fp32 FMA tiles, 1 tcgen05 for every shape it takes, 2 tcgen05 where it measured faster per call); ``None`` leaves the
    library's setting alone.  ``dense`` / ``cross_pair`` take the engine as an argument and re-apply it in their backward."""

    def __init__(self, engine):
        self.engine = engine

    def __enter__(self):
        self.prev = None if self.engine is None else _lib._lib.xdr_set_dense_engine(int(self.engine))
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            _lib._lib.xdr_set_dense_engine(self.prev)
        return False


class Dense(torch.autograd.Function):
    """``act(X W^T + b + mask * (X2 W2^T))``: nn.Linear + activation (emcdr.py:86-93, recbole MLPLayers dtcdr.py:61-67)
    and, with X2/W2/mask, one CoNet cross-stitch unit (conet.py:118-138, mask = ids < n_overlap, conet.py:113-116)."""

    @staticmethod
    def forward(ctx, X, W, bias, X2, W2, mask_ids, mask_lt, act, engine=None):
        X = X.contiguous()
        _require_cuda_f32(X, 'X')
        _require_cuda_f32(W, 'W')
        M, K = X.shape
        N = W.shape[0]
        if W.shape[1] != K:
            raise ValueError(f'dense: X is [*, {K}] but W is {tuple(W.shape)}')
        if X2 is not None:
            X2 = X2.contiguous()
            if X2.shape != X.shape or W2.shape != W.shape:
                raise ValueError('dense: cross operands must have the shapes of X and W')
        Y = torch.empty((M, N), dtype=torch.float32, device=X.device)
        with dense_engine(engine):
            call('xdr_dense_fwd', ptr(X), ptr(W), ptr(bias), ptr(X2), ptr(W2), ptr(mask_ids), int(mask_lt), int(act), ptr(Y), M,
                 N, K, cur_stream())
        ctx.save_for_backward(X, W, bias, X2, W2, mask_ids, Y)
        ctx.mask_lt, ctx.act, ctx.engine = int(mask_lt), int(act), engine
        return Y

    @staticmethod
    def backward(ctx, dY):
        with dense_engine(ctx.engine):
            return Dense._backward(ctx, dY)

    @staticmethod
    def _backward(ctx, dY):
        X, W, bias, X2, W2, mask_ids, Y = ctx.saved_tensors
        M, K = X.shape
        N = W.shape[0]
        s = cur_stream()
        dY = dY.contiguous()
        dZ = torch.empty_like(dY)
        call('xdr_act_bwd', ptr(Y), ptr(dY), ctx.act, ptr(dZ), dY.numel(), s)
        needs = ctx.needs_input_grad
        dX = dW = db = dX2 = dW2 = None
        # the weight-gradient kernels accumulate: their destinations (dW, db, dW2) are views of ONE zero-filled block -- one
        # fill kernel per layer instead of three
        want_w = needs[1] or (bias is not None and needs[2])
        want_w2 = X2 is not None and needs[4]
        pad4 = lambda n: (n + 3) // 4 * 4   # every view starts 16-byte aligned
        sizes = [pad4(W.numel()) if want_w else 0, pad4(bias.numel()) if (want_w and bias is not None) else 0,
                 pad4(W2.numel()) if want_w2 else 0]
        if sum(sizes):
            parts = torch.split(torch.zeros(sum(sizes), dtype=torch.float32, device=X.device), sizes)
        if needs[0]:
            dX = torch.empty_like(X)
            call('xdr_dense_bwd_input', ptr(dZ), ptr(W), None, 0, ptr(dX), M, N, K, 0, s)
        if want_w:
            dW = parts[0][:W.numel()].view_as(W)
            db = parts[1][:bias.numel()].view_as(bias) if bias is not None else None
            call('xdr_dense_bwd_weight', ptr(dZ), ptr(X), None, 0, ptr(dW), ptr(db), M, N, K, s)
        if X2 is not None:
            if needs[3]:
                dX2 = torch.empty_like(X2)
                call('xdr_dense_bwd_input', ptr(dZ), ptr(W2), ptr(mask_ids), ctx.mask_lt, ptr(dX2), M, N, K, 0, s)
            if want_w2:
                dW2 = parts[2][:W2.numel()].view_as(W2)
                call('xdr_dense_bwd_weight', ptr(dZ), ptr(X2), ptr(mask_ids), ctx.mask_lt, ptr(dW2), None, M, N, K, s)
        return dX, dW, db, dX2, dW2, None, None, None, None


def dense(X, W, bias=None, act=_lib.ACT_NONE, X2=None, W2=None, mask_ids=None, mask_lt=0, engine=None):
    return Dense.apply(X, W, bias, X2, W2, mask_ids, mask_lt, act, engine)


# Independent launches of one autograd node on parallel streams (fork from the current stream, join back into it; inside a
# CUDA-graph capture the lanes become parallel branches of the graph).  Only LAUNCHES go to the side streams -- every tensor is
# allocated on the current stream before the fork and used after the join, so the caching allocator's per-stream pools never see
# a cross-stream free.  -1 (default) = two lanes inside a CUDA-graph capture (trainer.GraphedTrainStep: CoNet's BOTH step at
# BASELINE configs[2] 995 -> 900 us on a B200, profiles/r2_conet_stacked.md), none in eager steps (the fork / join events are
# host work an eager step cannot hide); 0 = never; n > 1 = always n lanes.
CROSS_STREAMS = int(_os.environ.get('XDR_CROSS_STREAMS', '-1'))
_SIDE_STREAMS = {}


def set_cross_streams(n: int) -> int:
    """Streams ``cross_pair`` spreads its independent launches over: -1 two inside a graph capture, else none; 0 / 1 none;
    n > 1 always n.  Returns the previous setting."""
    global CROSS_STREAMS
    prev, CROSS_STREAMS = CROSS_STREAMS, max(-1, int(n))
    return prev


def side_streams(device, n: int):
    """The cached side streams of ``device`` (created on first use; trainer.GraphedTrainStep asks for them before it starts a
    capture so that no stream is created inside one)."""
    sides = _SIDE_STREAMS.setdefault(_lib._device_key(device), [])
    while len(sides) < n:
        sides.append(torch.cuda.Stream(device=device))
    return sides[:n]


def _run_lanes(device, lanes):
    """Runs the callables of ``lanes``: lane 0 (mod n) on the current stream, the others on side streams."""
    n = CROSS_STREAMS
    if n < 0:
        n = 2 if (device.type == 'cuda' and torch.cuda.is_current_stream_capturing()) else 0
    n = min(n, len(lanes))
    if n <= 1 or device.type != 'cuda':
        for lane in lanes:
            lane()
        return
    main = torch.cuda.current_stream(device)
    sides = side_streams(device, n - 1)
    used = []
    for i, lane in enumerate(lanes):
        k = i % n
        if k == 0:
            continue
        st = sides[k - 1]
        if st not in used:
            st.wait_stream(main)
            used.append(st)
        with torch.cuda.stream(st):
            lane()
    for i, lane in enumerate(lanes):
        if i % n == 0:
            lane()
    for st in used:
        main.wait_stream(st)


class CrossPair(torch.autograd.Function):
    """Both directions of one CoNet cross-stitch layer (conet.py:118-138) as ONE autograd node:
    ``h_s = act(x_s Ws^T + bs + m * (x_t H^T))``, ``h_t = act(x_t Wt^T + bt + m * (x_s H^T))`` with the SAME ``H`` in both.
    Two ``Dense`` nodes would leave autograd to add the two halves of d x_s, d x_t and dH with element-wise kernels and to
    zero-fill two blocks; here the second input-gradient product accumulates into the first one's result
    (``xdr_dense_bwd_input(accumulate=1)``), both dH products add into one destination, and every weight gradient of the
    layer lives in one zero-filled block: 11 launches per layer backward instead of 15."""

    @staticmethod
    def forward(ctx, x_s, x_t, Ws, bs, Wt, bt, H, mask_ids, mask_lt, act, engine=None):
        x_s, x_t = x_s.contiguous(), x_t.contiguous()
        for t, nm in ((x_s, 'x_s'), (x_t, 'x_t'), (Ws, 'Ws'), (Wt, 'Wt'), (H, 'H')):
            _require_cuda_f32(t, nm)
        M, K = x_s.shape
        N = Ws.shape[0]
        if x_t.shape != x_s.shape or Ws.shape != (N, K) or Wt.shape != (N, K) or H.shape != (N, K):
            raise ValueError(f'cross_pair: x {tuple(x_s.shape)} / {tuple(x_t.shape)}, weights {tuple(Ws.shape)} / '
                             f'{tuple(Wt.shape)} / {tuple(H.shape)}')
        h_s = torch.empty((M, N), dtype=torch.float32, device=x_s.device)
        h_t = torch.empty((M, N), dtype=torch.float32, device=x_s.device)
        with dense_engine(engine):
            _run_lanes(x_s.device, [
                lambda: call('xdr_dense_fwd', ptr(x_s), ptr(Ws), ptr(bs), ptr(x_t), ptr(H), ptr(mask_ids), int(mask_lt), int(act),
                             ptr(h_s), M, N, K, cur_stream()),
                lambda: call('xdr_dense_fwd', ptr(x_t), ptr(Wt), ptr(bt), ptr(x_s), ptr(H), ptr(mask_ids), int(mask_lt), int(act),
                             ptr(h_t), M, N, K, cur_stream())])
        ctx.save_for_backward(x_s, x_t, Ws, bs, Wt, bt, H, mask_ids, h_s, h_t)
        ctx.mask_lt, ctx.act, ctx.engine = int(mask_lt), int(act), engine
        ctx.set_materialize_grads(False)
        return h_s, h_t

    @staticmethod
    def backward(ctx, d_hs, d_ht):
        with dense_engine(ctx.engine):
            return CrossPair._backward(ctx, d_hs, d_ht)

    @staticmethod
    def _backward(ctx, d_hs, d_ht):
        x_s, x_t, Ws, bs, Wt, bt, H, mask_ids, h_s, h_t = ctx.saved_tensors
        M, K = x_s.shape
        N = Ws.shape[0]
        s = cur_stream()
        needs = ctx.needs_input_grad

        def pre_act(h, dh):   # d(pre-activation); None when that output took no part in the loss
            if dh is None:
                return None
            dh = dh.contiguous()
            dz = torch.empty_like(dh)
            call('xdr_act_bwd', ptr(h), ptr(dh), ctx.act, ptr(dz), dh.numel(), s)
            return dz

        dz_s, dz_t = pre_act(h_s, d_hs), pre_act(h_t, d_ht)
        pad4 = lambda n: (n + 3) // 4 * 4   # every view starts 16-byte aligned
        want = [needs[2] or (bs is not None and needs[3]), needs[4] or (bt is not None and needs[5]), needs[6]]
        sizes = [pad4(Ws.numel()) if want[0] else 0, pad4(bs.numel()) if (want[0] and bs is not None) else 0,
                 pad4(Wt.numel()) if want[1] else 0, pad4(bt.numel()) if (want[1] and bt is not None) else 0,
                 pad4(H.numel()) if want[2] else 0]
        if sum(sizes):
            parts = torch.split(torch.zeros(sum(sizes), dtype=torch.float32, device=x_s.device), sizes)

        # every destination is allocated here, on the current stream; the launches below may run on parallel streams (_run_lanes)
        def d_input(dz_own, W_own, dz_other):   # dX = dz_own W_own + m * (dz_other H); returns (dX, its launches)
            if dz_own is None and dz_other is None:
                return None, None
            dX = torch.empty((M, K), dtype=torch.float32, device=x_s.device)

            def lane():
                if dz_own is not None:
                    call('xdr_dense_bwd_input', ptr(dz_own), ptr(W_own), None, 0, ptr(dX), M, N, K, 0, cur_stream())
                if dz_other is not None:
                    call('xdr_dense_bwd_input', ptr(dz_other), ptr(H), ptr(mask_ids), ctx.mask_lt, ptr(dX), M, N, K,
                         0 if dz_own is None else 1, cur_stream())
            return dX, lane

        d_xs, lane_xs = d_input(dz_s, Ws, dz_t) if needs[0] else (None, None)
        d_xt, lane_xt = d_input(dz_t, Wt, dz_s) if needs[1] else (None, None)
        dWs = dbs = dWt = dbt = dH = None
        if want[0]:
            dWs = parts[0][:Ws.numel()].view_as(Ws)
            dbs = parts[1][:bs.numel()].view_as(bs) if bs is not None else None
        if want[1]:
            dWt = parts[2][:Wt.numel()].view_as(Wt)
            dbt = parts[3][:bt.numel()].view_as(bt) if bt is not None else None
        if want[2]:
            dH = parts[4][:H.numel()].view_as(H)

        def lane_w():    # the towers' own weights (and biases)
            if want[0] and dz_s is not None:
                call('xdr_dense_bwd_weight', ptr(dz_s), ptr(x_s), None, 0, ptr(dWs), ptr(dbs), M, N, K, cur_stream())
            if want[1] and dz_t is not None:
                call('xdr_dense_bwd_weight', ptr(dz_t), ptr(x_t), None, 0, ptr(dWt), ptr(dbt), M, N, K, cur_stream())

        def lane_h():    # the shared cross-stitch matrix: both products add into one destination
            if want[2] and dz_s is not None:
                call('xdr_dense_bwd_weight', ptr(dz_s), ptr(x_t), ptr(mask_ids), ctx.mask_lt, ptr(dH), None, M, N, K, cur_stream())
            if want[2] and dz_t is not None:
                call('xdr_dense_bwd_weight', ptr(dz_t), ptr(x_s), ptr(mask_ids), ctx.mask_lt, ptr(dH), None, M, N, K, cur_stream())

        _run_lanes(x_s.device, [ln for ln in (lane_xs, lane_xt, lane_w, lane_h) if ln is not None])
        return d_xs, d_xt, dWs, dbs, dWt, dbt, dH, None, None, None, None


def cross_pair(x_s, x_t, Ws, bs, Wt, bt, H, mask_ids, mask_lt, act=_lib.ACT_RELU, engine=None):
    return CrossPair.apply(x_s, x_t, Ws, bs, Wt, bt, H, mask_ids, mask_lt, act, engine)


class MseRows(torch.autograd.Function):
    """``nn.MSELoss(Y, tgt_tab[idx])`` with the target embedding NOT detached (EMCDR.calculate_map_loss,
    emcdr.py:156-168): gradient flows to Y and, as a scatter-add, to the target table."""

    @staticmethod
    def forward(ctx, Y, tgt_tab, idx):
        Y = Y.contiguous()
        _require_cuda_f32(Y, 'Y')
        _require_cuda_f32(tgt_tab, 'target table')
        idx = _ids(idx, 'idx').reshape(-1)
        dev = Y.device
        out8 = torch.empty(8, dtype=torch.float32, device=dev)
        call('xdr_mse_rows_fwd', ptr(Y), ptr(tgt_tab), tgt_tab.shape[0], tgt_tab.shape[1], ptr(idx), idx.numel(), ptr(out8),
             ptr(_lib.workspace(dev)), _oob(dev), cur_stream())
        _maybe_check(dev)
        ctx.save_for_backward(Y, tgt_tab, idx)
        return out8[0]

    @staticmethod
    def backward(ctx, grad_loss):
        Y, tgt_tab, idx = ctx.saved_tensors
        g = grad_loss.reshape(-1)[:1].contiguous().float()
        dY = torch.empty_like(Y)
        dt, rt = _grad_dst(tgt_tab)
        call('xdr_mse_rows_bwd', ptr(Y), ptr(tgt_tab), tgt_tab.shape[0], tgt_tab.shape[1], ptr(idx), idx.numel(), ptr(g), 1.0,
             ptr(dY), ptr(dt), cur_stream())
        return dY, rt, None


def mse_rows(Y, tgt_tab, idx):
    return MseRows.apply(Y, tgt_tab, idx)


class BceLogit(torch.autograd.Function):
    """``nn.BCELoss(sigmoid(logit), label)`` (conet.py:140,196-197; dtcdr.py:121-124,186-187); returns (loss, prob)."""

    @staticmethod
    def forward(ctx, logit, label):
        logit = logit.contiguous().reshape(-1)
        _require_cuda_f32(logit, 'logit')
        _require_cuda_f32(label, 'label')
        dev = logit.device
        prob = torch.empty_like(logit)
        out8 = torch.empty(8, dtype=torch.float32, device=dev)
        call('xdr_bce_logit_fwd', ptr(logit), ptr(label), logit.numel(), ptr(prob), ptr(out8), ptr(_lib.workspace(dev)),
             cur_stream())
        ctx.save_for_backward(prob, label)
        ctx.mark_non_differentiable(prob)
        return out8[0], prob

    @staticmethod
    def backward(ctx, grad_loss, _grad_prob):
        prob, label = ctx.saved_tensors
        g = grad_loss.reshape(-1)[:1].contiguous().float()
        dlogit = torch.empty_like(prob)
        call('xdr_bce_logit_bwd', ptr(prob), ptr(label), prob.numel(), ptr(g), ptr(dlogit), cur_stream())
        return dlogit, None


def bce_logit(logit, label):
    return BceLogit.apply(logit, label)


class FrobSum(torch.autograd.Function):
    """``sum(torch.norm(H) for H in mats)`` -- CoNet's regulariser over the cross-stitch matrices (conet.py:198-201) -- as one
    launch forward and one backward instead of a reduction + add per matrix and four element-wise kernels per matrix."""
    MAX_MATS = 8

    @staticmethod
    def forward(ctx, *mats):
        for i, m in enumerate(mats):
            _require_cuda_f32(m, f'matrix {i}')
        dev = mats[0].device
        buf = torch.empty(len(mats) + 1, dtype=torch.float32, device=dev)    # [norms..., sum]
        counts = (_ct.c_int64 * len(mats))(*[m.numel() for m in mats])
        call('xdr_frob_sum_fwd', _ptr_array(mats), counts, len(mats), buf.data_ptr(), buf[len(mats):].data_ptr(), cur_stream())
        ctx.save_for_backward(buf, *mats)
        return buf[len(mats)]

    @staticmethod
    def backward(ctx, grad_loss):
        buf, *mats = ctx.saved_tensors
        g = grad_loss.reshape(-1)[:1].contiguous().float()
        dsts = [torch.empty_like(m) for m in mats]
        counts = (_ct.c_int64 * len(mats))(*[m.numel() for m in mats])
        call('xdr_frob_sum_bwd', _ptr_array(mats), counts, len(mats), buf.data_ptr(), ptr(g), _ptr_array(dsts), cur_stream())
        return tuple(dsts)


def frob_sum(mats):
    """Sum of the Frobenius norms of ``mats`` (contiguous fp32 tensors; at most ``FrobSum.MAX_MATS``)."""
    return FrobSum.apply(*mats)


def select_dot(mapped, tgt_tab, sel_ids, n_overlap, other_tab, other_ids):
    """EMCDR.predict tail (emcdr.py:191-205): where(id < n_overlap, mapped, Et[id]) . other[id2]; inference only."""
    with torch.no_grad():
        sel_ids, other_ids = _ids(sel_ids, 'sel_ids'), _ids(other_ids, 'other_ids')
        b, dev = sel_ids.numel(), tgt_tab.device
        score = torch.empty(b, dtype=torch.float32, device=dev)
        call('xdr_select_dot', ptr(mapped.contiguous()), ptr(tgt_tab), tgt_tab.shape[0], ptr(sel_ids), int(n_overlap),
             ptr(other_tab), other_tab.shape[0], ptr(other_ids), tgt_tab.shape[1], b, ptr(score), _oob(dev), cur_stream())
        _maybe_check(dev)
        return score


def mlp_chain(x, weights: Sequence[torch.Tensor], biases: Sequence[Optional[torch.Tensor]], hidden_act: int,
              last_act: int):
    """Linear -> act -> ... -> Linear -> last_act, every layer through xdr_dense_fwd."""
    n = len(weights)
    for k in range(n):
        x = dense(x, weights[k], biases[k], last_act if k == n - 1 else hidden_act)
    return x


# ------------------------------------------------------------------------------------------------------------------
# A17: K training steps in one persistent launch
# ------------------------------------------------------------------------------------------------------------------

_steps_ws = {}


def _steps_workspace(device, n_steps):
    index = device.index if device.index is not None else (torch.cuda.current_device() if device.type == 'cuda' else -1)
    key = (index, cur_stream())
    need = _lib._lib.xdr_steps_workspace_bytes(int(n_steps))
    ws = _steps_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _steps_ws[key] = ws
    return ws


def set_steps_early_scatter(on: bool):
    """Opt in to the early-scatter form of the persistent kernel for launches with ``reg_weight == 0`` (CMF's yaml default
    lambda = gamma = 0, reg-free BPR): the row gradients do not depend on the batch-wide EmbLoss norms then, so the scatter
    warps do not wait for the step's norm exchange.  Results are the same (tests/test_gpu_engines.py on a B200); opt-in."""
    _lib._lib.xdr_steps_set_early_scatter(1 if on else 0)


def train_steps_supported(batch: int, dim: int, pairwise: bool, device=None) -> bool:
    """True when (batch, dim) is a shape the persistent kernels take (mirror of plan_steps in steps_persistent.cu)."""
    if batch <= 0 or batch % 4 != 0 or dim % 4 != 0 or dim > 256:
        return False
    props = torch.cuda.get_device_properties(device if device is not None else torch.cuda.current_device())
    sms = props.multi_processor_count
    nv = dim // 4
    lpr = 8 if nv <= 16 else (16 if nv <= 32 else 32)
    rows = 3 if pairwise else 2
    slice_ = max(4, (-(-batch // sms) + 3) // 4 * 4)
    grid = -(-batch // slice_)
    if grid > sms or grid > 160:
        return False
    tasks = -(-slice_ // (32 // lpr))
    up = lambda v: (v + 127) // 128 * 128
    base = up(384 + 8 * tasks * 16) + 8 * up(rows * slice_ * 8 + slice_ * 4)
    stage = up(rows * slice_ * dim * 4 + 2 * slice_ * 4)
    if base + 3 * stage <= 220 * 1024:   # staged kernel (3 or 4 stages)
        return True
    return tasks <= 2 * 20 and base <= 220 * 1024  # register kernel


_hot_rows_keepalive = {}


def set_steps_hot_rows(hot_users=None, hot_items=None):
    """Hot rows of the following ``train_steps`` launches (``xdr_steps_set_hot_rows``): int64 device tensors of row ids whose
    gradients every CTA pre-aggregates in shared memory (popular items of a Zipf-like catalogue).  ``None`` / empty switches
    a side off.  The tensors are kept alive here until replaced."""
    hu = hot_users.contiguous() if hot_users is not None and hot_users.numel() else None
    hi = hot_items.contiguous() if hot_items is not None and hot_items.numel() else None
    for t in (hu, hi):
        if t is not None and (t.dtype != torch.int64 or not _on_device(t)):
            raise ValueError('hot row ids must be int64 device tensors')
    _hot_rows_keepalive['u'], _hot_rows_keepalive['i'] = hu, hi
    call('xdr_steps_set_hot_rows', ptr(hu), 0 if hu is None else hu.numel(), ptr(hi), 0 if hi is None else hi.numel())


def hot_rows_from_ids(ids, n_rows: int, k: int = 32, min_share: float = 2e-3):
    """The (at most ``k``) rows that ``ids`` names most often and that each take at least ``min_share`` of all occurrences --
    a dataset statistic (popular items), computed once.  Returns an int64 tensor (possibly empty)."""
    flat = ids.reshape(-1)
    cnt = torch.bincount(flat, minlength=int(n_rows))
    top = torch.topk(cnt, min(k, cnt.numel()))
    keep = top.values.float() >= min_share * float(flat.numel())
    return top.indices[keep].to(torch.int64)


class TouchMap:
    """Which rows of a pair of gradient tables hold gradients -- the state of *lazily zeroed* gradient tables
    (``xdr_train_steps_lazy``, include/xdr.h).  2 bits per row (bit 0 claimed, bit 1 zero-filled), user part first.  A row
    whose bits are clear counts as zero whatever the table holds; ``clear()`` is therefore ``zero_grad()`` at N/4 bytes
    instead of the reference's dense N x D fill."""

    def __init__(self, n_users: int, n_items: int, device):
        self.n_users, self.n_items = int(n_users), int(n_items)
        nbytes = _lib._lib.xdr_touch_map_bytes(self.n_users, self.n_items)
        self.words = torch.zeros(nbytes // 4, dtype=torch.int32, device=device)
        self._user_words = ((self.n_users + 15) // 16 + 3) // 4 * 4

    def clear(self):
        self.words.zero_()

    def _mask(self, words, n):
        bits = (words.view(-1, 1) >> (2 * torch.arange(16, device=words.device, dtype=torch.int32))) & 3
        return bits.reshape(-1)[:n]

    def state(self):
        """Per-row 2-bit states ``(users [n_users], items [n_items])``: 0 untouched, 3 claimed and zero-filled."""
        return self._mask(self.words[:self._user_words], self.n_users), self._mask(self.words[self._user_words:], self.n_items)

    def touched(self):
        """Boolean masks ``(users, items)`` of the rows that hold gradients."""
        su, si = self.state()
        return su != 0, si != 0


def train_steps(user_tab, item_tab, user, item_a, item_b=None, label=None, *, loss_kind=_lib.LOSS_MSE, reg_weight=0.0,
                gamma=1e-10, user_dst=None, item_dst=None, scale=1.0, grad_loss=None, out8=None, touch=None, fresh=False):
    """Run ``K = user.shape[0]`` training steps (fwd + bwd + scatter-add) in ONE persistent launch.

    ``user`` / ``item_a`` / ``item_b`` (pairwise) / ``label`` (pointwise) are ``[K, B]`` device tensors (row k = batch
    k; rows may be strided views of a larger ``[K, ..]`` buffer as long as each row is contiguous).  Gradients are
    scatter-added into ``user_dst`` / ``item_dst`` scaled by ``scale`` (default: dense gradient tables ``.grad``-style;
    pass the weight tables themselves and ``scale=-lr`` for fused asynchronous SGD).  Returns ``out8`` ``[K, 8]`` whose
    column 0 is the per-step loss -- the value ``calculate_loss`` returns for that batch.
    ``touch`` (a ``TouchMap``): the destinations are lazily zeroed gradient tables -- rows the map does not mark count as
    zero and are zero-filled on first touch, so the scatter-adds never read gradient lines from DRAM; ``fresh=True`` clears
    the map first (the tables then hold the gradient of exactly these K batches on the marked rows).
    Replaces K iterations of recbole ``Trainer._train_epoch`` around emcdr.py:110-154 / cmf.py:75-98.
    """
    _require_cuda_f32(user_tab, 'user table')
    _require_cuda_f32(item_tab, 'item table')
    pairwise = item_b is not None
    K, B = user.shape
    for t, nm in ((user, 'user'), (item_a, 'item_a'), (item_b, 'item_b')):
        if t is None:
            continue
        if t.dtype != torch.int64 or not _on_device(t) or t.shape != (K, B) or t.stride(1) != 1:
            raise ValueError(f'{nm} must be a CUDA int64 [K, B] tensor with contiguous rows')
        if t.stride(0) != user.stride(0):
            raise ValueError('all id tensors must share the same step stride')
    if label is not None and (label.dtype != torch.float32 or label.shape != (K, B) or label.stride(0) != user.stride(0)):
        raise ValueError('label must be float32 [K, B] with the step stride of the id tensors')
    dev = user_tab.device
    if user_dst is None:
        user_dst = torch.zeros_like(user_tab)
    if item_dst is None:
        item_dst = torch.zeros_like(item_tab)
    if out8 is None:
        out8 = torch.empty((K, 8), dtype=torch.float32, device=dev)
    ws = _steps_workspace(dev, K)
    if touch is not None:
        if touch.n_users != user_tab.shape[0] or touch.n_items != item_tab.shape[0]:
            raise ValueError('touch map was built for other table sizes')
        call('xdr_train_steps_lazy', ptr(user_tab), ptr(item_tab), user_tab.shape[0], item_tab.shape[0], user_tab.shape[1],
             ptr(user), ptr(item_a), ptr(item_b), ptr(label), user.stride(0), B, K, 1 if pairwise else 0, int(loss_kind),
             float(gamma), float(reg_weight), ptr(grad_loss), float(scale), ptr(user_dst), ptr(item_dst), ptr(out8), ptr(ws),
             ws.numel(), ptr(touch.words), 1 if fresh else 0, _oob(dev), cur_stream())
    else:
        call('xdr_train_steps', ptr(user_tab), ptr(item_tab), user_tab.shape[0], item_tab.shape[0], user_tab.shape[1],
             ptr(user), ptr(item_a), ptr(item_b), ptr(label), user.stride(0), B, K, 1 if pairwise else 0, int(loss_kind),
             float(gamma), float(reg_weight), ptr(grad_loss), float(scale), ptr(user_dst), ptr(item_dst), ptr(out8), ptr(ws),
             ws.numel(), _oob(dev), cur_stream())
    _maybe_check(dev)
    return out8, user_dst, item_dst


# ------------------------------------------------------------------------------------------------------------------
# A4 / A14-A15 fused: gather -> small MLP -> loss head -> backward -> scatter in one kernel
# ------------------------------------------------------------------------------------------------------------------
# fp32 FMA row tiles (fused_mlp.cu) / 3xTF32 mma.sync row tiles (tc_mlp.cu) / bf16x3 tcgen05 tiles with TMEM accumulators
# (tc5_mlp.cu: the EMCDR map stack [D, 128, D] only)
FUSED_MLP_ENGINES = ('fma', 'tc', 'tc5')
_FUSED_MLP_ENTRY = {'fma': 'xdr_fused_mlp', 'tc': 'xdr_tc_mlp', 'tc5': 'xdr_tc5_mlp'}


def fused_mlp_engine(flag) -> Optional[str]:
    """Engine named by the model config key ``xdr_fused_mlp``: False/None -> composed kernels, True/'fma' -> the fp32
    row-tile kernel, 'tc' -> the mma.sync tensor-core row-tile kernel, 'tc5' -> the tcgen05 kernel (EMCDR map stack only)."""
    if flag in (None, False, 0, '', 'false', 'False'):
        return None
    if flag in (True, 1, 'fma', 'true', 'True'):
        return 'fma'
    if flag in ('tc', 'tc5'):
        return flag
    raise ValueError(f"xdr_fused_mlp must be False, True, 'fma', 'tc' or 'tc5', got {flag!r}")


def fused_mlp_supported(dims, engine: str = 'fma') -> bool:
    arr = (_ct.c_int * len(dims))(*[int(d) for d in dims])
    fn = getattr(_lib._lib, _FUSED_MLP_ENTRY[engine] + '_supported')
    return bool(fn(len(dims) - 1, _ct.cast(arr, _ct.c_void_p)))


def _ptr_array(ts):
    return (_ct.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])


def _fused_mlp_call(in_mode, head, hidden_act, tabs, idx_u, idx_i, label, Ws, bs, backward, grad_loss, dsts, dWs, dbs,
                    want_prob, engine='fma'):
    Au, Bu, Ai, Bi, T = tabs
    dev = Au.device
    dims = [Ws[0].shape[1]] + [w.shape[0] for w in Ws]
    arr = (_ct.c_int * len(dims))(*dims)
    B = idx_u.numel()
    out8 = torch.empty(8, dtype=torch.float32, device=dev)
    prob = torch.empty(B, dtype=torch.float32, device=dev) if want_prob else None
    dAu, dBu, dAi, dBi, dT = dsts
    if engine not in FUSED_MLP_ENGINES:
        raise ValueError(f'unknown fused-MLP engine {engine!r}')
    call(_FUSED_MLP_ENTRY[engine] + '_step', len(Ws), _ct.cast(arr, _ct.c_void_p), _ptr_array(Ws), _ptr_array(bs),
         _ptr_array(dWs) if dWs else None, _ptr_array(dbs) if dbs else None, int(hidden_act), int(in_mode), int(head),
         ptr(Au), ptr(Bu), ptr(Ai), ptr(Bi), ptr(T), Au.shape[0], Ai.shape[0] if Ai is not None else 0, Au.shape[1],
         ptr(idx_u), ptr(idx_i), ptr(label), B, int(backward), ptr(grad_loss), 1.0, ptr(dAu), ptr(dBu), ptr(dAi),
         ptr(dBi), ptr(dT), ptr(prob), ptr(out8), ptr(_lib.workspace(dev)), _oob(dev), cur_stream())
    _maybe_check(dev)
    return out8, prob


class FusedMlpLoss(torch.autograd.Function):
    """One fused kernel for ``loss(MLP(gathered rows))`` and one for its whole backward (the forward is recomputed on
    chip; rows come back from L2).  in_mode 0 / head 0: EMCDR.calculate_map_loss (emcdr.py:156-168); in_mode 1 / head 1:
    one DTCDR NeuMF term (dtcdr.py:112-125, 186-187).  Tensor arguments: Au, Bu, Ai, Bi, T, then the layer weights, then
    the layer biases (None where a layer has none)."""

    @staticmethod
    def forward(ctx, in_mode, head, hidden_act, idx_u, idx_i, label, n_layers, engine, eager, Au, Bu, Ai, Bi, T, *wb):
        Ws, bs = list(wb[:n_layers]), list(wb[n_layers:])
        for t in (Au, Bu, Ai, Bi, T) + tuple(Ws):
            if t is not None:
                _require_cuda_f32(t, 'fused_mlp operand')
        idx_u = _ids(idx_u, 'idx_u').reshape(-1)
        idx_i = _ids(idx_i, 'idx_i').reshape(-1) if idx_i is not None else None
        ctx.cfg = (in_mode, head, hidden_act, n_layers, engine)
        ctx.eager = None
        if eager:
            # ONE launch for the loss and every gradient (upstream gradient 1, what loss.backward() passes): table gradients go
            # straight into the tables' .grad ('inplace' mode), the MLP's into one zeroed scratch block that backward() returns;
            # backward() adds the (g - 1)-fold with a correction launch that is over at once when g == 1 (xdr.h)
            dsts = [None if t is None else _grad_dst(t)[0] for t in (Au, Bu, Ai, Bi, T)]
            sizes = [w.numel() for w in Ws] + [0 if b is None else b.numel() for b in bs]
            flat = torch.zeros(sum(sizes), dtype=torch.float32, device=Au.device)
            parts = list(torch.split(flat, sizes))
            dWs = [parts[k].view_as(w) for k, w in enumerate(Ws)]
            dbs = [None if b is None else parts[n_layers + k].view_as(b) for k, b in enumerate(bs)]
            out8, _ = _fused_mlp_call(in_mode, head, hidden_act, (Au, Bu, Ai, Bi, T), idx_u, idx_i, label, Ws, bs, 1, None, dsts,
                                      dWs, dbs, False, engine)
            ctx.eager = (dsts, dWs, dbs)
        else:
            out8, _ = _fused_mlp_call(in_mode, head, hidden_act, (Au, Bu, Ai, Bi, T), idx_u, idx_i, label, Ws, bs, 0, None,
                                      (None,) * 5, None, None, False, engine)
        ctx.save_for_backward(idx_u, idx_i, label, Au, Bu, Ai, Bi, T, *Ws, *bs)
        return out8[0]

    @staticmethod
    def backward(ctx, grad_loss):
        in_mode, head, hidden_act, n_layers, engine = ctx.cfg
        sv = ctx.saved_tensors
        idx_u, idx_i, label, Au, Bu, Ai, Bi, T = sv[:8]
        Ws, bs = list(sv[8:8 + n_layers]), list(sv[8 + n_layers:])
        g = grad_loss.reshape(-1)[:1].contiguous().float()
        if ctx.eager is not None:
            dsts, dWs, dbs = ctx.eager
            ctx.eager = None   # (the returned gradients are then unshared: autograd adopts them as .grad instead of copying)
            _fused_mlp_call(in_mode, head, hidden_act, (Au, Bu, Ai, Bi, T), idx_u, idx_i, label, Ws, bs, 2, g, dsts, dWs, dbs,
                            False, engine)
            return (None,) * 9 + (None,) * 5 + tuple(dWs) + tuple(dbs)
        dsts, rets = [], []
        for t in (Au, Bu, Ai, Bi, T):
            if t is None:
                dsts.append(None)
                rets.append(None)
            else:
                d, r = _grad_dst(t)
                dsts.append(d)
                rets.append(r)
        dWs = [torch.zeros_like(w) for w in Ws]
        dbs = [None if b is None else torch.zeros_like(b) for b in bs]
        _fused_mlp_call(in_mode, head, hidden_act, (Au, Bu, Ai, Bi, T), idx_u, idx_i, label, Ws, bs, 1, g, dsts, dWs, dbs,
                        False, engine)
        return (None,) * 9 + tuple(rets) + tuple(dWs) + tuple(dbs)


def fused_mlp_loss(in_mode, head, hidden_act, idx_u, idx_i, label, tabs, Ws, bs, engine='fma'):
    """``loss(MLP(gathered rows))`` with its whole backward in the fused kernels.  With the tcgen05 engine and 'inplace' table
    gradients the step is EAGER: forward computes the loss and accumulates every gradient in one launch, backward only
    corrects for an upstream gradient other than 1 (see FusedMlpLoss.forward).  Consequence of eagerness, as for every
    in-place gradient: a forward under grad mode whose backward is never called has still added to the tables' ``.grad``."""
    Au, Bu, Ai, Bi, T = tabs
    tables = [t for t in tabs if t is not None]
    eager = (engine == 'tc5' and EAGER_TC5 and _TABLE_GRAD_MODE == 'inplace' and torch.is_grad_enabled() and
             all(t.is_leaf and t.requires_grad for t in tables) and all(w.requires_grad for w in Ws))
    return FusedMlpLoss.apply(in_mode, head, hidden_act, idx_u, idx_i, label, len(Ws), engine, eager, Au, Bu, Ai, Bi, T, *Ws, *bs)


def fused_mlp_prob(hidden_act, idx_u, idx_i, tabs, Ws, bs, engine='fma'):
    """Forward only, sigmoid output per row (DTCDR.predict)."""
    with torch.no_grad():
        idx_u, idx_i = _ids(idx_u, 'idx_u').reshape(-1), _ids(idx_i, 'idx_i').reshape(-1)
        label = torch.zeros(idx_u.numel(), dtype=torch.float32, device=tabs[0].device)
        _, prob = _fused_mlp_call(1, 1, hidden_act, tabs, idx_u, idx_i, label, list(Ws), list(bs), False, None, (None,) * 5,
                                  None, None, True, engine)
        return prob


# ------------------------------------------------------------------------------------------------------------------
# A7-A8 fused: one CoNet tower pass (cross-stitch stack + BCE + backward + scatter) in one tensor-core kernel
# ------------------------------------------------------------------------------------------------------------------

def conet_fused_supported(dims, dim) -> bool:
    arr = (_ct.c_int * len(dims))(*[int(d) for d in dims])
    return bool(_lib._lib.xdr_tc_conet_supported(len(dims) - 1, _ct.cast(arr, _ct.c_void_p), int(dim)))


def _conet_call(want, mask_on_item, n_overlap, user, item, label, tabs, w_out, b_out, ws, bs, wt, bt, hs, backward,
                grad_loss, dsts, grads, want_prob):
    Su, Si, Tu, Ti = tabs
    dev = Su.device
    dims = [ws[0].shape[1]] + [w.shape[0] for w in ws]
    arr = (_ct.c_int * len(dims))(*dims)
    B = user.numel()
    out8 = torch.empty(8, dtype=torch.float32, device=dev)
    prob = torch.empty(B, dtype=torch.float32, device=dev) if want_prob else None
    scratch = torch.empty((B, 2 * dims[1]), dtype=torch.float32, device=dev) if backward else None
    dSu, dSi, dTu, dTi = dsts
    g = grads if grads is not None else dict(ws=None, bs=None, wt=None, bt=None, h=None, w_out=None, b_out=None)
    pa = lambda lst: None if lst is None else _ptr_array(lst)
    call('xdr_tc_conet_step', len(ws), _ct.cast(arr, _ct.c_void_p), _ptr_array(ws), _ptr_array(bs), _ptr_array(wt),
         _ptr_array(bt), _ptr_array(hs), pa(g['ws']), pa(g['bs']), pa(g['wt']), pa(g['bt']), pa(g['h']), ptr(w_out), ptr(b_out),
         ptr(g['w_out']), ptr(g['b_out']), int(want), ptr(Su), ptr(Si), ptr(Tu), ptr(Ti), Su.shape[0], Si.shape[0], Su.shape[1],
         ptr(user), ptr(item), ptr(label), B, 1 if mask_on_item else 0, int(n_overlap), 1 if backward else 0, ptr(grad_loss), 1.0,
         ptr(dSu), ptr(dSi), ptr(dTu), ptr(dTi), ptr(scratch), ptr(prob), ptr(out8), ptr(_lib.workspace(dev)), _oob(dev),
         cur_stream())
    _maybe_check(dev)
    return out8, prob


class ConetTowerLoss(torch.autograd.Function):
    """``BCELoss(source_forward(u, i), label)`` (want 0) or ``BCELoss(target_forward(u, i), label)`` (want 1) of CoNet
    (conet.py:105-181, 196-197) as ONE kernel forward and ONE kernel for the whole backward (the forward is recomputed on
    chip).  Tensor arguments: the four tables, the wanted output unit (weight, bias), then per layer the source weights,
    source biases, target weights, target biases, cross parameters (5 * n_layers tensors)."""

    @staticmethod
    def forward(ctx, want, mask_on_item, n_overlap, user, item, label, n_layers, Su, Si, Tu, Ti, w_out, b_out, *params):
        L = n_layers
        ws, bs, wt, bt, hs = (list(params[k * L:(k + 1) * L]) for k in range(5))
        for t in (Su, Si, Tu, Ti, w_out, b_out) + tuple(params):
            _require_cuda_f32(t, 'conet operand')
        _require_cuda_f32(label, 'label')
        user, item = _ids(user, 'user').reshape(-1), _ids(item, 'item').reshape(-1)
        out8, _ = _conet_call(want, mask_on_item, n_overlap, user, item, label, (Su, Si, Tu, Ti), w_out, b_out, ws, bs, wt, bt,
                              hs, False, None, (None,) * 4, None, False)
        ctx.cfg = (want, mask_on_item, n_overlap, L)
        ctx.save_for_backward(user, item, label, Su, Si, Tu, Ti, w_out, b_out, *params)
        return out8[0]

    @staticmethod
    def backward(ctx, grad_loss):
        want, mask_on_item, n_overlap, L = ctx.cfg
        sv = ctx.saved_tensors
        user, item, label, Su, Si, Tu, Ti, w_out, b_out = sv[:9]
        params = sv[9:]
        ws, bs, wt, bt, hs = (list(params[k * L:(k + 1) * L]) for k in range(5))
        g = grad_loss.reshape(-1)[:1].contiguous().float()
        dsts, rets = [], []
        for t in (Su, Si, Tu, Ti):
            d, r = _grad_dst(t)
            dsts.append(d)
            rets.append(r)
        grads = dict(ws=[torch.zeros_like(w) for w in ws], bs=[torch.zeros_like(b) for b in bs],
                     wt=[torch.zeros_like(w) for w in wt], bt=[torch.zeros_like(b) for b in bt],
                     h=[torch.zeros_like(h) for h in hs], w_out=torch.zeros_like(w_out), b_out=torch.zeros_like(b_out))
        _conet_call(want, mask_on_item, n_overlap, user, item, label, (Su, Si, Tu, Ti), w_out, b_out, ws, bs, wt, bt, hs, True, g,
                    dsts, grads, False)
        return (None,) * 7 + tuple(rets) + (grads['w_out'], grads['b_out']) + tuple(grads['ws']) + tuple(grads['bs']) + \
            tuple(grads['wt']) + tuple(grads['bt']) + tuple(grads['h'])


def conet_tower_loss(want, mask_on_item, n_overlap, user, item, label, tabs, w_out, b_out, ws, bs, wt, bt, hs):
    Su, Si, Tu, Ti = tabs
    return ConetTowerLoss.apply(want, mask_on_item, n_overlap, user, item, label, len(ws), Su, Si, Tu, Ti, w_out, b_out,
                                *ws, *bs, *wt, *bt, *hs)


# ------------------------------------------------------------------------------------------------------------------
# F1: row-sparse optimizer step (SGD / Adagrad / lazy Adam) over the rows a batch touched
# ------------------------------------------------------------------------------------------------------------------

def sparse_optim_rows(kind, table, grad, ids, stamp, step_id, lr, *, state1=None, state2=None, adam_t=1, eps=1e-8,
                      beta1=0.9, beta2=0.999):
    """In place: for every distinct id, update ``table[id]`` (and the state rows) from ``grad[id]`` and zero ``grad[id]``.
    Replaces ``optimizer.zero_grad()`` + ``optimizer.step()`` of recbole ``Trainer._train_epoch`` for an embedding table
    (torch.optim.SGD / Adagrad / SparseAdam semantics, see xdr.h).  ``stamp``: int32 ``[n_rows]`` zeros at start;
    ``step_id`` >= 1 and strictly increasing per call on the same ``stamp``."""
    _require_cuda_f32(table, 'table')
    _require_cuda_f32(grad, 'grad table')
    if grad.shape != table.shape:
        raise ValueError('grad table must have the shape of the weight table')
    for st, nm in ((state1, 'state1'), (state2, 'state2')):
        if st is not None:
            _require_cuda_f32(st, nm)
            if st.shape != table.shape:
                raise ValueError(f'{nm} must have the shape of the weight table')
    if stamp.dtype != torch.int32 or stamp.numel() != table.shape[0] or not stamp.is_contiguous():
        raise ValueError('stamp must be a contiguous int32 [n_rows] tensor')
    ids = _ids(ids, 'ids').reshape(-1)
    call('xdr_sparse_optim_rows', int(kind), ptr(table), ptr(grad), ptr(state1), ptr(state2), ptr(stamp), ptr(ids),
         ids.numel(), table.shape[0], table.shape[1], int(step_id), int(adam_t), float(lr), float(eps), float(beta1),
         float(beta2), _oob(table.device), cur_stream())
    _maybe_check(table.device)
    return table


# ------------------------------------------------------------------------------------------------------------------
# F2: full-sort scoring + history mask + top-k without the [B, n_items] score matrix
# ------------------------------------------------------------------------------------------------------------------

def full_sort_topk(user_vecs, item_tab, k, n_items=None, first_item=1, hist_ptr=None, hist_ids=None, engine='mma'):
    """Top-``k`` items per user by ``user_vecs @ item_tab[first_item:n_items].T`` with each user's history excluded.

    The fused form of ``full_sort_predict`` (emcdr.py:208-233, cmf.py:107-112) + recbole's full-sort masking + ``topk``.
    ``hist_ptr`` ``[B + 1]`` / ``hist_ids``: int64 CSR of ascending item ids per user.  Returns ``(scores [B, k] fp32,
    ids [B, k] int64)``, score descending, ties by ascending id, padded with ``(-inf, -1)``.  ``engine``: ``'mma'``
    (mma.sync row tiles) or ``'tc5'`` (tcgen05.mma with tensor-memory accumulators; dim <= 64; on a B200 25 ms against 70 ms for ``'mma'`` at 4096 users x 1M
    items, profiles/r2_call1_new_kernels.jsonl)."""
    with torch.no_grad():
        user_vecs = user_vecs.contiguous()
        _require_cuda_f32(user_vecs, 'user_vecs')
        _require_cuda_f32(item_tab, 'item table')
        B, D = user_vecs.shape
        if item_tab.shape[1] != D:
            raise ValueError('user vectors and item rows must have the same width')
        n_items = item_tab.shape[0] if n_items is None else int(n_items)
        if hist_ptr is not None:
            hist_ptr, hist_ids = _ids(hist_ptr, 'hist_ptr'), _ids(hist_ids, 'hist_ids')
            if hist_ptr.numel() != B + 1:
                raise ValueError('hist_ptr must have batch + 1 entries')
        dev = user_vecs.device
        out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((B, k), dtype=torch.int64, device=dev)
        nbytes = _lib._lib.xdr_topk_workspace_bytes(B, int(k))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        if engine not in ('mma', 'tc5'):
            raise ValueError(f'unknown top-k engine {engine!r}')
        call('xdr_full_sort_topk_tc5' if engine == 'tc5' else 'xdr_full_sort_topk', ptr(user_vecs), B, ptr(item_tab), n_items, D, int(first_item), ptr(hist_ptr), ptr(hist_ids),
             int(k), ptr(out_s), ptr(out_i), ptr(ws), nbytes, cur_stream())
        return out_s, out_i
