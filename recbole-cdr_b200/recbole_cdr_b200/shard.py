"""Row-sharded embedding tables over peer memory -- the multi-GPU form of the hot path.

The reference is single-device; this is the B200 addition.  One process per GPU (``torch.distributed``; NCCL or gloo is
only used to exchange 64-byte CUDA-IPC handles and for barriers).  A table of ``n_rows`` global rows is split
block-cyclically: global row ``r`` lives on rank ``r % G`` at local row ``r // G`` -- that spreads the three id ranges of
the joint layout (overlapped / target-only / source-only, data/dataset.py:344-445) and Zipf-hot low ids evenly, whereas a
block split would put every overlapped row on rank 0.

Every rank maps every peer's shard (``xdr_ipc_open``) and hands all G pointers to ``xdr_train_steps_sharded``: the warps
that score a batch issue the remote ``LDG`` gathers and ``RED`` scatter-adds themselves, so the NVLink transfers overlap
the math interaction by interaction and there is no all-to-all step on the data path.

The batch is data-parallel: every rank trains on its own batches.  For NVLink locality the loader routes an interaction
to the rank that owns its user row (``user % G``): one of the three rows is then local and NVLink traffic drops by 1/3.
"""
import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, cur_stream, ptr


def owner_of(rows: torch.Tensor, world: int) -> torch.Tensor:
    return rows % world


def local_row_of(rows: torch.Tensor, world: int) -> torch.Tensor:
    return rows // world


def shard_rows(n_rows: int, world: int) -> int:
    """Rows per shard (the same on every rank: the last shards are padded with dead rows)."""
    return -(-n_rows // world)


def route_by_user_owner(user: torch.Tensor, world: int) -> List[torch.Tensor]:
    """Positions of a batch that belong to each rank under user-owner routing (``user % world``)."""
    own = owner_of(user, world)
    return [torch.nonzero(own == r, as_tuple=False).reshape(-1) for r in range(world)]


class RowShardedTable:
    """This rank's shard of a block-cyclically row-sharded ``[n_rows, dim]`` fp32 table + the peers' mapped shards."""

    def __init__(self, n_rows: int, dim: int, rank: int, world: int, device, local: Optional[torch.Tensor] = None):
        if world < 1 or world > 8 or world & (world - 1):
            raise ValueError('world size must be a power of two <= 8')
        self.n_rows, self.dim, self.rank, self.world = n_rows, dim, rank, world
        self.device = torch.device(device)
        rows = shard_rows(n_rows, world)
        if local is None:
            local = torch.zeros((rows, dim), dtype=torch.float32, device=self.device)
        if tuple(local.shape) != (rows, dim) or local.dtype != torch.float32 or not local.is_contiguous():
            raise ValueError(f'local shard must be a contiguous float32 [{rows}, {dim}] tensor')
        self.local = local
        self._ptrs = None       # device pointers of all shards as seen from this rank
        self._opened = []       # peer allocation bases to close

    # ---- construction helpers ------------------------------------------------------------------------------------
    @classmethod
    def from_full(cls, full: torch.Tensor, rank: int, world: int, device):
        """Take rows ``rank::world`` of a full table (tests / small tables)."""
        n_rows, dim = full.shape
        rows = shard_rows(n_rows, world)
        local = torch.zeros((rows, dim), dtype=torch.float32, device=device)
        mine = full[rank::world]
        local[:mine.shape[0]].copy_(mine)
        return cls(n_rows, dim, rank, world, device, local)

    def to_full(self, group=None) -> torch.Tensor:
        """All-gather the shards back into the global row order (tests / checkpoints)."""
        parts = [torch.empty_like(self.local) for _ in range(self.world)] if self.world > 1 else [self.local]
        if self.world > 1:
            dist.all_gather(parts, self.local, group=group)
        full = torch.empty((shard_rows(self.n_rows, self.world) * self.world, self.dim), dtype=torch.float32,
                           device=self.local.device)
        for r, p in enumerate(parts):
            full[r::self.world].copy_(p)
        return full[:self.n_rows]

    # ---- peer mapping --------------------------------------------------------------------------------------------
    def connect(self, group=None):
        """Exchange CUDA-IPC handles and map every peer's shard; must be called by all ranks."""
        if self.world == 1:
            self._ptrs = [self.local.data_ptr()]
            return self
        handle = (ctypes.c_ubyte * 64)()
        offset = ctypes.c_int64(0)
        call('xdr_ipc_export', self.local.data_ptr(), ctypes.cast(handle, ctypes.c_void_p), ctypes.byref(offset))
        mine = (bytes(handle), int(offset.value))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        ptrs = []
        for r, (h, off) in enumerate(everyone):
            if r == self.rank:
                ptrs.append(self.local.data_ptr())
                continue
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            base = ctypes.c_void_p(0)
            call('xdr_ipc_open', ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(base))
            self._opened.append(base.value)
            ptrs.append(base.value + off)
        self._ptrs = ptrs
        return self

    def pointer_array(self):
        if self._ptrs is None:
            raise RuntimeError('RowShardedTable.connect() has not been called')
        if getattr(self, '_ptr_array', None) is None:
            self._ptr_array = (ctypes.c_void_p * self.world)(*self._ptrs)   # built once: the launches are host-bound
        return self._ptr_array

    def rows_view(self, first_local_row: int, n_local_rows: int):
        """Local rows [first, first + n) of every shard as a block-cyclic table of its own (``n * world`` global rows):
        e.g. the user block and the item block of a node table."""
        return _RowsView(self, first_local_row, n_local_rows)

    def close(self):
        for base in self._opened:
            call('xdr_ipc_close', base)
        self._opened = []
        self._ptrs = None
        self._ptr_array = None


class _RowsView(object):
    def __init__(self, table: RowShardedTable, first: int, n: int):
        if first < 0 or n < 0 or first + n > table.local.shape[0]:
            raise ValueError('rows_view outside the shard')
        self.world, self.dim, self.rank = table.world, table.dim, table.rank
        self.n_rows = n * table.world
        self.local = table.local[first:first + n]
        off = first * table.dim * 4
        self._ptr_array = (ctypes.c_void_p * table.world)(*[p + off for p in table.pointer_array()])

    def pointer_array(self):
        return self._ptr_array


_steps_ws = {}


def train_steps_sharded(user_tab: RowShardedTable, item_tab: RowShardedTable, user_dst: RowShardedTable,
                        item_dst: RowShardedTable, user, item_a, item_b=None, label=None, *, loss_kind=_lib.LOSS_MSE,
                        reg_weight=0.0, gamma=1e-10, scale=1.0, out8=None, staged_a=None, staged_b=None):
    """K training steps of this rank's batches against the row-sharded tables in ONE persistent launch
    (``xdr_train_steps_sharded``); ids are GLOBAL row ids.  Arguments as in ``ops.train_steps``.  ``staged_a`` /
    ``staged_b``: optional dense ``[K, B, D]`` blocks holding the (first / second) item row of every interaction, pulled
    ahead by ``gather_rows_sharded``; without them the kernel gathers item rows straight from the (peer) shards."""
    K, B = user.shape
    dev = user_tab.local.device
    if out8 is None:
        out8 = torch.empty((K, 8), dtype=torch.float32, device=dev)
    need = _lib._lib.xdr_steps_workspace_bytes(int(K))
    key = (dev.index, cur_stream())
    ws = _steps_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _steps_ws[key] = ws
    call('xdr_train_steps_sharded', user_tab.pointer_array(), item_tab.pointer_array(), user_dst.pointer_array(),
         item_dst.pointer_array(), user_tab.world, user_tab.n_rows, item_tab.n_rows, user_tab.dim, ptr(user), ptr(item_a),
         ptr(item_b), ptr(label), user.stride(0), B, K, 1 if item_b is not None else 0, int(loss_kind), float(gamma),
         float(reg_weight), None, float(scale), ptr(out8), ptr(ws), ws.numel(), ptr(staged_a), ptr(staged_b), None, cur_stream())
    return out8


def gather_rows_sharded(table: RowShardedTable, idx: torch.Tensor, out: torch.Tensor):
    """out[k, :] = table[idx[k], :] with rows pulled from whichever rank owns them (peer ``LDG`` over NVLink).
    ``idx`` is flat, or a ``[K, B]`` view with contiguous rows (e.g. ``ids[:, 1]`` of a ``[K, 3, B]`` block)."""
    if idx.dim() == 2 and idx.stride(1) == 1:
        n, batch, stride = idx.numel(), idx.shape[1], idx.stride(0)
    else:
        idx = idx.reshape(-1).contiguous()
        n, batch, stride = idx.numel(), 0, 0
    call('xdr_gather_rows_sharded', table.pointer_array(), table.world, table.n_rows, table.dim, ptr(idx), n, batch, stride,
         ptr(out), table.dim, None, cur_stream())
    return out


class ShardedStepRunner(object):
    """Chunked, double-buffered driver of the sharded persistent kernel.

    Default (``stage_remote=False``): one persistent launch whose warps gather from and scatter-add into the peer shards
    directly.  ``stage_remote=True`` splits the remote side off: a massively parallel peer-gather kernel runs ONE CHUNK
    AHEAD on a second stream and pulls the chunk's item rows into dense local blocks that the persistent kernel then
    reads sequentially (only its gradient ``RED``s cross NVLink).  Measured in round 1 the staged variant is SLOWER at
    every GPU count (2: 7.8 vs 6.4, 4: 12.8 vs 10.5, 8: 14.5 vs 12.5 us/step): the staging block costs ~8 MB/step of
    local DRAM traffic and random 256-byte peer reads top out near 300 GB/s per GPU whoever issues them.  Kept as a
    tested option.  ``run(ids)``: ids ``[K, 3, B]`` (or ``[K, 2, B]`` + label) on the device; returns out8 ``[K, 8]``.
    """

    def __init__(self, user_tab, item_tab, user_dst, item_dst, *, pairwise=True, loss_kind=_lib.LOSS_MSE, reg_weight=0.0,
                 gamma=1e-10, scale=1.0, chunk=50, stage_remote=False):
        self.t = (user_tab, item_tab, user_dst, item_dst)
        self.kw = dict(loss_kind=loss_kind, reg_weight=reg_weight, gamma=gamma, scale=scale)
        self.pairwise, self.chunk = pairwise, int(chunk)
        self.stage_remote = stage_remote and user_tab.world > 1
        self.dev = user_tab.local.device
        self.fetch_stream = torch.cuda.Stream(device=self.dev)
        self._stage = None
        self._fetched = [torch.cuda.Event(), torch.cuda.Event()]
        self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self._used = [False, False]
        self.launches = 0

    def _buffers(self, B):
        D = self.t[1].dim
        n = 2 if self.pairwise else 1
        if self._stage is None or self._stage.shape[2:] != (self.chunk, B, D) or self._stage.shape[1] != n:
            self._stage = torch.empty((2, n, self.chunk, B, D), dtype=torch.float32, device=self.dev)
        return self._stage

    def _fetch(self, ids, c0, c1, slot):
        """peer-gather the item rows of steps [c0, c1) into staging buffer `slot` on the fetch stream"""
        st = self._buffers(ids.shape[2])
        with torch.cuda.stream(self.fetch_stream):
            if self._used[slot]:
                self.fetch_stream.wait_event(self._consumed[slot])
            for w in range(st.shape[1]):
                gather_rows_sharded(self.t[1], ids[c0:c1, 1 + w], st[slot, w])
            self._fetched[slot].record(self.fetch_stream)

    def run(self, ids: torch.Tensor, label=None, out8=None):
        K, R, B = ids.shape
        if out8 is None:
            out8 = torch.empty((K, 8), dtype=torch.float32, device=self.dev)
        main = torch.cuda.current_stream(self.dev)
        if not self.stage_remote:
            train_steps_sharded(*self.t, ids[:, 0], ids[:, 1], ids[:, 2] if self.pairwise else None, label, out8=out8, **self.kw)
            self.launches += 1
            return out8
        self.fetch_stream.wait_stream(main)          # the ids are produced on the main stream
        bounds = [(c, min(c + self.chunk, K)) for c in range(0, K, self.chunk)]
        self._fetch(ids, *bounds[0], 0)
        for n, (c0, c1) in enumerate(bounds):
            slot = n % 2
            if n + 1 < len(bounds):
                self._fetch(ids, *bounds[n + 1], (n + 1) % 2)   # overlaps with the persistent kernel of chunk n
            main.wait_event(self._fetched[slot])
            st = self._stage
            train_steps_sharded(*self.t, ids[c0:c1, 0], ids[c0:c1, 1], ids[c0:c1, 2] if self.pairwise else None,
                                None if label is None else label[c0:c1], out8=out8[c0:c1],
                                staged_a=st[slot, 0], staged_b=st[slot, 1] if self.pairwise else None, **self.kw)
            self._consumed[slot].record(main)
            self._used[slot] = True
            self.launches += 1 + st.shape[1]
        return out8
