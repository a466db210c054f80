"""BiTGCF over G GPUs: row-sharded graph propagation.

The reference is single-device (bitgcf.py:174-205 propagates the full graph of both domains for every batch); this is the
B200 multi-GPU form of the same arithmetic.  One process per GPU.  The joint node space of a domain (users, then items,
bitgcf.py:92-116) is padded so that the item block starts at a multiple of G and is then split block-cyclically like the
embedding tables of ``shard.py``: padded node ``v`` lives on rank ``v % G`` at local row ``v // G``; local rows are
``[ceil(Nu/G) users | ceil(Ni/G) items]``.  Because the overlapped users (items) are the ids below ``n_ov``,
they form a PREFIX of every rank's local user (item) block, so ``transfer_layer`` + ``F.normalize`` and the layer
combine stay row-local and run through the unchanged single-GPU kernels with local counts.

Only the sparse matmul needs remote rows: every rank holds the CSR rows of ``L`` it owns (global column ids); ``L`` is
symmetric, so the backward pass is the same launch on the gradient.  One exchange per layer and direction, two forms:

  exchange='allgather' (default)  the operand shards are all-gathered (NCCL, large coalesced NVLink messages; the target
                                  domain's gather overlaps the source domain's SpMM) into a local ``[G, rows, D]`` replica
                                  and ``xdr_spmm_csr_sharded`` reads it with the shard addressing.  With ~14 nonzeros per
                                  row nearly every row is needed by every rank, so moving each row once beats fetching it
                                  per nonzero.
  exchange='peer'                 ``xdr_spmm_csr_sharded`` gathers the neighbour rows straight from the peers' exchange
                                  buffers (CUDA-IPC mapped) -- no replica memory, but random 256-byte NVLink reads
                                  (~300 GB/s per GPU measured); a tiny all-reduce on the stream orders "owner wrote its
                                  shard" before "peers read it" (and the reverse before the buffer is reused).

The batch is data-parallel: the gathers from the propagated tables, sigmoid(dot) + BCE and the EmbLoss on the ego rows
(bitgcf.py:207-250) are one ``xdr_train_steps_sharded`` launch per domain and term, whose gradient rows land in the
owners' gradient shards by peer ``RED``; every rank then back-propagates its local shard of that gradient.  The
objective is the mean over ranks of the reference's per-batch loss (the usual data-parallel convention).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, cur_stream, ptr
from .graph import NormAdj, TransferNorm, _elementwise, norm_adj_coo
from .shard import RowShardedTable, train_steps_sharded


class NodeShards(object):
    """Padded joint node space of one rank: users ``[0, Ru*G)``, items ``[Ru*G, (Ru+Ri)*G)``."""

    def __init__(self, n_users, n_items, n_ov_users, n_ov_items, rank, world):
        self.n_users, self.n_items, self.rank, self.world = n_users, n_items, rank, world
        self.ru, self.ri = -(-n_users // world), -(-n_items // world)
        self.rows = self.ru + self.ri                   # local rows
        self.n_padded = self.rows * world
        # overlapped ids are [0, n_ov): locally a prefix of ceil((n_ov - rank) / G) rows
        self.ov_users = max(0, -(-(n_ov_users - rank) // world))
        self.ov_items = max(0, -(-(n_ov_items - rank) // world))

    def pad(self, v):
        """joint node id (user u -> u, item i -> n_users + i)  ->  padded node id"""
        v = np.asarray(v, dtype=np.int64)
        return np.where(v < self.n_users, v, v - self.n_users + self.ru * self.world)

    def local_nodes(self):
        """joint node id held by every local row, -1 for padding rows"""
        g = np.arange(self.rows, dtype=np.int64)
        u = g[:self.ru] * self.world + self.rank
        i = g[:self.ri] * self.world + self.rank
        return np.concatenate([np.where(u < self.n_users, u, -1), np.where(i < self.n_items, i + self.n_users, -1)])

    def take_local(self, node_values: torch.Tensor) -> torch.Tensor:
        """rows of a full ``[n_users + n_items, ...]`` tensor that this rank owns (padding rows zero)"""
        nodes = torch.from_numpy(self.local_nodes()).to(node_values.device)
        out = node_values[nodes.clamp_min(0)].clone()
        out[nodes < 0] = 0
        return out


class ShardedNormAdj(NormAdj):
    """This rank's rows of ``L`` (``NormAdj``) with padded global column ids: identical values and per-row order."""

    def __init__(self, rows, cols, nodes: NodeShards, device, chunk=256):
        r, c, val = norm_adj_coo(rows, cols, nodes.n_users, nodes.n_items)     # global degrees -> global values
        rp, cp = nodes.pad(r), nodes.pad(c)
        mine = (rp % nodes.world) == nodes.rank
        self.n, self.n_users, self.n_items = nodes.n_padded, nodes.n_users, nodes.n_items
        self.device = torch.device(device)
        self.nodes = nodes
        self._set_csr(rp[mine] // nodes.world, cp[mine], val[mine], nodes.rows, chunk)  # r sorted => local rows sorted

    def spmm_sharded(self, x: RowShardedTable):
        """S = (this rank's rows of L) . X with X row-sharded over the peers"""
        S = torch.empty_like(x.local)
        call('xdr_spmm_csr_sharded', ptr(self.work_row), ptr(self.work_beg), ptr(self.work_end), ptr(self.work_split),
             self.work_row.numel(), ptr(self.split_rows), self.split_rows.numel(), ptr(self.col), ptr(self.val),
             x.pointer_array(), x.world, x.dim, ptr(S), cur_stream())
        return S


class _Replica(object):
    """a local ``[G, rows, D]`` all-gathered copy of a row-sharded operand, addressed like the shards themselves"""

    def __init__(self, world, rows, dim, device):
        import ctypes
        self.world, self.dim = world, dim
        self.buf = torch.empty((world, rows, dim), dtype=torch.float32, device=device)
        self.local = self.buf[0]
        self._ptrs = (ctypes.c_void_p * world)(*[self.buf[g].data_ptr() for g in range(world)])

    def pointer_array(self):
        return self._ptrs


class _ShardedProp(torch.autograd.Function):
    """``graph_layer`` of both domains (bitgcf.py:130-135, drop_rate 0) on the local rows; see ``graph.GraphProp``."""

    @staticmethod
    def forward(ctx, Es, Et, eng):
        Es, Et = Es.contiguous(), Et.contiguous()
        Ss, St = eng.spmm_pair(Es, Et)
        ctx.eng = eng
        ctx.save_for_backward(Es, Et, Ss, St)
        return _elementwise(Es, Ss, None, 0), _elementwise(Et, St, None, 0)

    @staticmethod
    def backward(ctx, Gs, Gt):
        Es, Et, Ss, St = ctx.saved_tensors
        Gs = torch.zeros_like(Es) if Gs is None else Gs.contiguous()
        Gt = torch.zeros_like(Et) if Gt is None else Gt.contiguous()
        Ts, Tt = ctx.eng.spmm_pair(_elementwise(Gs, Es, None, 1), _elementwise(Gt, Et, None, 1))
        return _elementwise(Gs, Ss, Ts, 2), _elementwise(Gt, St, Tt, 2), None


class ShardedBiTGCF(object):
    """BiTGCF training step over ``world`` GPUs.  ``src_edges`` / ``tgt_edges``: (user, item) arrays of each domain's
    interaction matrix in the joint id space; ``ego``: optional full tables (source_user, source_item, target_user,
    target_item) to take the local rows from (tests / checkpoints), else xavier-normal like the reference."""

    def __init__(self, src_edges, tgt_edges, n_users, n_items, n_ov_users, n_ov_items, *, dim, n_layers, lambda_source,
                 lambda_target, connect_way, reg_weight, rank, world, device, ego=None, group=None, exchange='allgather'):
        if exchange not in ('allgather', 'peer'):
            raise ValueError("exchange must be 'allgather' or 'peer'")
        self.exchange = exchange if world > 1 else 'peer'
        self.rank, self.world, self.device, self.group = rank, world, torch.device(device), group
        self.dim, self.n_layers, self.connect_way, self.reg_weight = dim, n_layers, connect_way, reg_weight
        self.lam_s, self.lam_t = float(lambda_source), float(lambda_target)
        nd = self.nodes = NodeShards(n_users, n_items, n_ov_users, n_ov_items, rank, world)
        self.adj_s = ShardedNormAdj(src_edges[0], src_edges[1], nd, device)
        self.adj_t = ShardedNormAdj(tgt_edges[0], tgt_edges[1], nd, device)

        def degrees(edges):  # bitgcf.py:79-82 with unit values: interactions per user / per item, duplicates counted
            du = np.bincount(np.asarray(edges[0], dtype=np.int64), minlength=n_users)
            di = np.bincount(np.asarray(edges[1], dtype=np.int64), minlength=n_items)
            return nd.take_local(torch.from_numpy(np.concatenate([du, di]).astype(np.float32))).to(self.device)
        self.deg_s, self.deg_t = degrees(src_edges), degrees(tgt_edges)

        self.out_dim = dim * (n_layers + 1) if connect_way == 'concat' else dim
        if connect_way not in ('concat', 'mean'):
            raise ValueError(f'connect_way [{connect_way}] is not supported')
        table = lambda d: RowShardedTable(nd.n_padded, d, rank, world, self.device)
        self.ego_s, self.ego_t = table(dim), table(dim)                # the embeddings (users then items)
        self.ego_grad_s, self.ego_grad_t = table(dim), table(dim)      # EmbLoss gradient rows arrive here by peer RED
        if self.exchange == 'peer':
            self.x_s, self.x_t = table(dim), table(dim)                # SpMM operand exchange buffers (peer-mapped)
            xbufs = [self.x_s, self.x_t]
        else:
            self.x_s, self.x_t = _Replica(world, nd.rows, dim, self.device), _Replica(world, nd.rows, dim, self.device)
            xbufs = []
        self.fin_s, self.fin_t = table(self.out_dim), table(self.out_dim)            # propagated tables
        self.fin_grad_s, self.fin_grad_t = table(self.out_dim), table(self.out_dim)  # their gradient
        if ego is not None:
            su, si, tu, ti = ego
            self.ego_s.local.copy_(nd.take_local(torch.cat([su, si], 0)))
            self.ego_t.local.copy_(nd.take_local(torch.cat([tu, ti], 0)))
        else:
            std = (2.0 / (n_users + dim)) ** 0.5   # xavier_normal_ of an [N, dim] embedding weight
            for t in (self.ego_s, self.ego_t):
                t.local.normal_(0.0, std)
                t.local[torch.from_numpy(nd.local_nodes() < 0).to(self.device)] = 0
        self._tables = [self.ego_s, self.ego_t, self.ego_grad_s, self.ego_grad_t, self.fin_s, self.fin_t, self.fin_grad_s,
                        self.fin_grad_t] + xbufs
        for t in self._tables:
            t.connect(group)
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.ego_s.local.requires_grad_(True)
        self.ego_t.local.requires_grad_(True)

    # ---- plumbing ------------------------------------------------------------------------------------------------
    def fence(self):
        """stream-ordered barrier over the ranks: work enqueued after it starts after every rank's earlier work is done"""
        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)

    def spmm_pair(self, Xs, Xt):
        if self.exchange == 'allgather':
            hs = dist.all_gather_into_tensor(self.x_s.buf, Xs.detach(), group=self.group, async_op=True)
            ht = dist.all_gather_into_tensor(self.x_t.buf, Xt.detach(), group=self.group, async_op=True)
            hs.wait()
            Ss = self.adj_s.spmm_sharded(self.x_s)
            ht.wait()                      # the target domain's gather ran under the source domain's SpMM
            return Ss, self.adj_t.spmm_sharded(self.x_t)
        self.fence()                       # the peers have finished reading the previous operands
        with torch.no_grad():
            self.x_s.local.copy_(Xs)
            self.x_t.local.copy_(Xt)
        self.fence()                       # every owner's operand shard is written
        return self.adj_s.spmm_sharded(self.x_s), self.adj_t.spmm_sharded(self.x_t)

    def _views(self, t: RowShardedTable):
        """(user rows, item rows) of a node table as block-cyclic tables of their own"""
        nd = self.nodes
        return t.rows_view(0, nd.ru), t.rows_view(nd.ru, nd.ri)

    # ---- the model -------------------------------------------------------------------------------------------------
    def forward(self):
        """bitgcf.py:174-205 on the local rows: returns this rank's rows of the combined source / target tables"""
        nd = self.nodes
        es, et = self.ego_s.local, self.ego_t.local
        ls, lt = [es], [et]
        for _ in range(self.n_layers):
            ps, pt = _ShardedProp.apply(es, et, self)
            es, et, ns, nt = TransferNorm.apply(ps, pt, self.deg_s, self.deg_t, nd.ru, nd.ri, nd.ov_users, nd.ov_items,
                                                self.lam_s, self.lam_t)
            ls.append(ns)
            lt.append(nt)
        if self.connect_way == 'concat':
            return torch.cat(ls, 1), torch.cat(lt, 1)
        def layer_mean(layers):   # running sum instead of stack + mean: half the table-sized passes (see BiTGCF._layer_mean)
            acc = layers[0]
            for e in layers[1:]:
                acc = acc + e
            return acc * (1.0 / len(layers))
        return layer_mean(ls), layer_mean(lt)

    def train_step(self, source_batch, target_batch):
        """``calculate_loss`` + ``backward`` (bitgcf.py:207-250) for this rank's batch: (user, item, label) per domain
        with GLOBAL ids.  Returns (source_loss, target_loss), each of shape [1], of THIS rank's batch; leaves
        d(mean over ranks of source_loss + target_loss)/d(ego rows) in ``ego_s.local.grad`` / ``ego_t.local.grad``."""
        fs, ft = self.forward()
        with torch.no_grad():
            self.fin_s.local.copy_(fs)
            self.fin_t.local.copy_(ft)
            for t in (self.fin_grad_s, self.fin_grad_t, self.ego_grad_s, self.ego_grad_t):
                t.local.zero_()
        self.fence()                       # propagated tables published, gradient shards cleared everywhere
        scale = 1.0 / self.world
        losses = []
        for (u, i, y), fin, fin_grad, ego, ego_grad in ((source_batch, self.fin_s, self.fin_grad_s, self.ego_s, self.ego_grad_s),
                                                        (target_batch, self.fin_t, self.fin_grad_t, self.ego_t, self.ego_grad_t)):
            u, i, y = u.reshape(1, -1), i.reshape(1, -1), y.reshape(1, -1).to(torch.float32)
            (fu, fi), (gu, gi) = self._views(fin), self._views(fin_grad)
            bce = train_steps_sharded(fu, fi, gu, gi, u, i, None, y, loss_kind=_lib.LOSS_BCE_SIGMOID, scale=scale)
            (eu, ei), (du, di) = self._views(ego), self._views(ego_grad)
            reg = train_steps_sharded(eu, ei, du, di, u, i, None, None, loss_kind=_lib.LOSS_NONE,
                                      reg_weight=self.reg_weight, scale=scale)
            losses.append(bce[0, 0:1] + reg[0, 0:1])
        self.fence()                       # every rank's gradient rows have landed in their owners' shards
        torch.autograd.backward([fs, ft], [self.fin_grad_s.local, self.fin_grad_t.local])
        with torch.no_grad():
            self.ego_s.local.grad.add_(self.ego_grad_s.local)
            self.ego_t.local.grad.add_(self.ego_grad_t.local)
        return tuple(losses)

    def full_tables(self, what='ego'):
        """all-gather node tables back into the reference's (users, items) layout (tests / checkpoints):
        'ego' -> the embeddings, 'grad' -> their gradient"""
        nd, out = self.nodes, []
        for t in ((self.ego_s, self.ego_t)):
            src = t.local.detach() if what == 'ego' else t.local.grad
            parts = [torch.empty_like(src) for _ in range(self.world)]
            if self.world > 1:
                dist.all_gather(parts, src.contiguous(), group=self.group)
            else:
                parts = [src]
            full = torch.empty((nd.n_padded, src.shape[1]), dtype=src.dtype, device=src.device)
            for r, p in enumerate(parts):
                full[r::self.world].copy_(p)
            out.append(full[:nd.n_users])
            out.append(full[nd.ru * self.world:nd.ru * self.world + nd.n_items])
        return out    # source_user, source_item, target_user, target_item

    def close(self):
        for t in self._tables:
            t.close()
