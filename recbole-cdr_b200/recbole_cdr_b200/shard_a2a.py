"""The all-to-all form of the row-sharded training step (as BASELINE.json's north_star words it:
"a single NCCL all-to-all of looked-up rows per batch").

Per batch and rank: bucket the batch's ids by owner -> all-to-all of the ids -> every owner gathers the requested rows from
its shard -> all-to-all of the rows back -> score + loss + row gradients on the received rows (the same per-step kernels as
on one GPU, run on a batch-sized "mini table" whose row p is the row of interaction p) -> all-to-all of the gradient rows ->
every owner scatter-adds them into its gradient (or weight) shard.  User and item requests travel in the same messages, so
a step is one size exchange + three data exchanges, whatever the number of tables.

This is the BASELINE form of the exchange: it moves the same bytes as the peer-memory kernel (``shard.train_steps_sharded``,
in-kernel remote loads / REDs over CUDA-IPC mappings, which is what ``bench.py`` runs) but pays a collective launch per
exchange and a host read of the message sizes per step, and nothing overlaps.  It exists so that the two can be compared on
the same hardware, and as the path that needs no peer mappings (any ``torch.distributed`` backend; gloo in the CPU tests).
Every kernel on it is one of the single-GPU kernels (gather, fused score + loss, scatter-add).

Equivalence (tests): per-rank losses equal the reference's per-batch loss on that rank's batch, and the re-assembled
gradient tables equal the dense autograd gradients over the union of the batches.
"""
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from .shard import RowShardedTable


def _bucket(ids: torch.Tensor, world: int):
    """Stable bucket sort of global row ids by owner (``id % world``): (order, local rows in bucket order, counts [world])."""
    owner = ids % world
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)
    return order, (ids // world)[order], counts


class AllToAllStep(object):
    """One training step over row-sharded tables with all-to-all exchanges (see the module docstring).

    ``user_tab`` / ``item_tab``: this rank's :class:`RowShardedTable` shards (no ``connect()`` needed);
    ``user_dst`` / ``item_dst``: shards that receive ``scale`` x the gradient rows (gradient shards, or the weight shards
    themselves with ``scale = -lr``).
    """

    def __init__(self, user_tab: RowShardedTable, item_tab: RowShardedTable, user_dst: RowShardedTable,
                 item_dst: RowShardedTable, *, pairwise=True, loss_kind=_lib.LOSS_MSE, reg_weight=0.0, gamma=1e-10, group=None):
        self.ut, self.it, self.du, self.di = user_tab, item_tab, user_dst, item_dst
        self.world, self.rank = user_tab.world, user_tab.rank
        self.pairwise, self.loss_kind, self.reg_weight, self.gamma = pairwise, loss_kind, reg_weight, gamma
        self.group = group
        self.exchanged_rows = 0     # rows this rank received as a requester so far (for traffic accounting)

    def _a2a(self, out, inp, out_splits, in_splits):
        if self.world == 1:
            out.copy_(inp)
        else:
            dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.group)
        return out

    def step(self, user: torch.Tensor, item_a: torch.Tensor, item_b: Optional[torch.Tensor] = None,
             label: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
        G, D, dev = self.world, self.ut.dim, user.device
        B = user.numel()
        items = torch.cat([item_a, item_b]) if self.pairwise else item_a
        # ---- 1. requests: [user rows wanted from r | item rows wanted from r] per owner r -----------------------------------
        order_u, rows_u, cnt_u = _bucket(user.reshape(-1), G)
        order_i, rows_i, cnt_i = _bucket(items.reshape(-1), G)
        send_cnt = torch.stack([cnt_u, cnt_i], dim=1)                       # [G, 2]
        recv_cnt = torch.empty_like(send_cnt)
        self._a2a(recv_cnt.view(-1), send_cnt.view(-1).contiguous(), None, None)
        send_cnt_h, recv_cnt_h = send_cnt.tolist(), recv_cnt.tolist()      # the one host read of the step (message sizes)
        in_splits = [a + b for a, b in send_cnt_h]
        out_splits = [a + b for a, b in recv_cnt_h]
        # interleave the two request lists per destination
        su, si = torch.split(rows_u, [c[0] for c in send_cnt_h]), torch.split(rows_i, [c[1] for c in send_cnt_h])
        req = torch.cat([t for pair in zip(su, si) for t in pair])
        got_req = torch.empty(sum(out_splits), dtype=torch.int64, device=dev)
        self._a2a(got_req, req, out_splits, in_splits)
        # ---- 2. owners gather the requested rows ------------------------------------------------------------------------------
        seg = torch.split(got_req, [c for pair in recv_cnt_h for c in pair])   # per source rank: user part, item part
        own_u = torch.cat(seg[0::2]) if G > 1 else seg[0]
        own_i = torch.cat(seg[1::2]) if G > 1 else seg[1]
        rows_for_u = ops.gather_rows_raw(self.ut.local, own_u)
        rows_for_i = ops.gather_rows_raw(self.it.local, own_i)
        ru, ri = torch.split(rows_for_u, [c[0] for c in recv_cnt_h]), torch.split(rows_for_i, [c[1] for c in recv_cnt_h])
        reply = torch.cat([t for pair in zip(ru, ri) for t in pair])
        got_rows = torch.empty((sum(in_splits), D), dtype=torch.float32, device=dev)
        self._a2a(got_rows, reply, [s for s in in_splits], [s for s in out_splits])
        self.exchanged_rows += got_rows.shape[0]
        # ---- 3. un-bucket into batch order: the mini tables ---------------------------------------------------------------------
        parts = torch.split(got_rows, [c for pair in send_cnt_h for c in pair])
        mini_u = torch.empty((B, D), dtype=torch.float32, device=dev)
        mini_i = torch.empty((items.numel(), D), dtype=torch.float32, device=dev)
        mini_u[order_u] = torch.cat(parts[0::2])
        mini_i[order_i] = torch.cat(parts[1::2])
        mini_u.requires_grad_(True)
        mini_i.requires_grad_(True)
        # ---- 4. score + loss + row gradients with the single-GPU kernels ----------------------------------------------------------
        pos = torch.arange(B, dtype=torch.int64, device=dev)
        mode = ops.get_table_grad_mode()
        ops.set_table_grad_mode('autograd')        # the mini tables' gradients are wanted as tensors, not added in place
        try:
            if self.pairwise:
                loss = ops.bpr_loss(mini_u, mini_i, pos, pos, pos + B, self.reg_weight, self.gamma)
            else:
                loss = ops.point_loss(mini_u, mini_i, pos, pos, label, self.loss_kind, self.reg_weight)
            gu, gi = torch.autograd.grad(loss.sum(), [mini_u, mini_i])
        finally:
            ops.set_table_grad_mode(mode)
        # ---- 5. gradient rows back to their owners, scatter-add ----------------------------------------------------------------------
        gu_b, gi_b = torch.split(gu[order_u], [c[0] for c in send_cnt_h]), torch.split(gi[order_i], [c[1] for c in send_cnt_h])
        send_g = torch.cat([t for pair in zip(gu_b, gi_b) for t in pair])
        got_g = torch.empty((sum(out_splits), D), dtype=torch.float32, device=dev)
        self._a2a(got_g, send_g, out_splits, in_splits)
        gseg = torch.split(got_g, [c for pair in recv_cnt_h for c in pair])
        ops.scatter_add_rows_raw(self.du.local, own_u, torch.cat(gseg[0::2]).contiguous(), scale)
        ops.scatter_add_rows_raw(self.di.local, own_i, torch.cat(gseg[1::2]).contiguous(), scale)
        return loss.detach()


class AllToAllChunkRunner(AllToAllStep):
    """K steps per exchange: the ids of a whole ``[K, R, B]`` block are bucketed at once, ONE round of all-to-alls brings every
    row the block needs, the persistent multi-step kernel (``ops.train_steps``) runs the K steps on the block-sized mini
    tables in a single launch, and one all-to-all returns the K steps' gradient rows to their owners.  Four collectives and
    one host read per BLOCK instead of per step, in messages K times larger (bulk NVLink transfers instead of per-row
    requests).  Exact in gradient-accumulation mode (the weights do not change inside a block); with ``scale = -lr`` on the
    weight shards the steps of a block read the weights as they were when the block started (the same staleness as the
    asynchronous fused-SGD launch).  ``run`` returns the ``[K]`` per-step losses of this rank's batches."""

    def run(self, ids: torch.Tensor, label: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
        G, D, dev = self.world, self.ut.dim, ids.device
        K, R, B = ids.shape
        if R != (3 if self.pairwise else 2):
            raise ValueError(f'id block must be [K, {3 if self.pairwise else 2}, B]')
        user = ids[:, 0].reshape(-1)                                           # position k*B + j
        items = ids[:, 1:].reshape(K, -1).reshape(-1)                          # position k*(R-1)*B + (r-1)*B + j
        order_u, rows_u, cnt_u = _bucket(user, G)
        order_i, rows_i, cnt_i = _bucket(items, G)
        send_cnt = torch.stack([cnt_u, cnt_i], dim=1)
        recv_cnt = torch.empty_like(send_cnt)
        self._a2a(recv_cnt.view(-1), send_cnt.view(-1).contiguous(), None, None)
        send_cnt_h, recv_cnt_h = send_cnt.tolist(), recv_cnt.tolist()          # the one host read of the block
        in_splits, out_splits = [a + b for a, b in send_cnt_h], [a + b for a, b in recv_cnt_h]
        su, si = torch.split(rows_u, [c[0] for c in send_cnt_h]), torch.split(rows_i, [c[1] for c in send_cnt_h])
        got_req = torch.empty(sum(out_splits), dtype=torch.int64, device=dev)
        self._a2a(got_req, torch.cat([t for pair in zip(su, si) for t in pair]), out_splits, in_splits)
        seg = torch.split(got_req, [c for pair in recv_cnt_h for c in pair])
        own_u, own_i = torch.cat(seg[0::2]), torch.cat(seg[1::2])
        ru = torch.split(ops.gather_rows_raw(self.ut.local, own_u), [c[0] for c in recv_cnt_h])
        ri = torch.split(ops.gather_rows_raw(self.it.local, own_i), [c[1] for c in recv_cnt_h])
        got_rows = torch.empty((sum(in_splits), D), dtype=torch.float32, device=dev)
        self._a2a(got_rows, torch.cat([t for pair in zip(ru, ri) for t in pair]), in_splits, out_splits)
        self.exchanged_rows += got_rows.shape[0]
        parts = torch.split(got_rows, [c for pair in send_cnt_h for c in pair])
        mini_u = torch.empty((K * B, D), dtype=torch.float32, device=dev)
        mini_i = torch.empty((K * (R - 1) * B, D), dtype=torch.float32, device=dev)
        mini_u[order_u] = torch.cat(parts[0::2])
        mini_i[order_i] = torch.cat(parts[1::2])
        # ---- K steps in one persistent launch on the mini tables: row of (step k, slot j) sits at a fixed position --------------
        j = torch.arange(B, dtype=torch.int64, device=dev)
        k0 = torch.arange(K, dtype=torch.int64, device=dev).view(-1, 1)
        pos = torch.stack([k0 * B + j] + [k0 * (R - 1) * B + r * B + j for r in range(R - 1)], dim=1).contiguous()   # [K, R, B]
        gu, gi = torch.zeros_like(mini_u), torch.zeros_like(mini_i)
        out8, _, _ = ops.train_steps(mini_u, mini_i, pos[:, 0], pos[:, 1], pos[:, 2] if self.pairwise else None, label,
                                     loss_kind=self.loss_kind, reg_weight=self.reg_weight, gamma=self.gamma, user_dst=gu,
                                     item_dst=gi, scale=1.0)
        # ---- gradient rows back to their owners ------------------------------------------------------------------------------------
        gu_b, gi_b = torch.split(gu[order_u], [c[0] for c in send_cnt_h]), torch.split(gi[order_i], [c[1] for c in send_cnt_h])
        got_g = torch.empty((sum(out_splits), D), dtype=torch.float32, device=dev)
        self._a2a(got_g, torch.cat([t for pair in zip(gu_b, gi_b) for t in pair]), out_splits, in_splits)
        gseg = torch.split(got_g, [c for pair in recv_cnt_h for c in pair])
        ops.scatter_add_rows_raw(self.du.local, own_u, torch.cat(gseg[0::2]).contiguous(), scale)
        ops.scatter_add_rows_raw(self.di.local, own_i, torch.cat(gseg[1::2]).contiguous(), scale)
        return out8[:, 0]
