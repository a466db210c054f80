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
        return (ctypes.c_void_p * self.world)(*self._ptrs)

    def close(self):
        for base in self._opened:
            call('xdr_ipc_close', base)
        self._opened = []
        self._ptrs = None


_steps_ws = {}


def train_steps_sharded(user_tab: RowShardedTable, item_tab: RowShardedTable, user_dst: RowShardedTable,
                        item_dst: RowShardedTable, user, item_a, item_b=None, label=None, *, loss_kind=_lib.LOSS_MSE,
                        reg_weight=0.0, gamma=1e-10, scale=1.0, out8=None):
    """K training steps of this rank's batches against the row-sharded tables in ONE persistent launch
    (``xdr_train_steps_sharded``); ids are GLOBAL row ids.  Arguments as in ``ops.train_steps``."""
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
         float(reg_weight), None, float(scale), ptr(out8), ptr(ws), ws.numel(), None, cur_stream())
    return out8
