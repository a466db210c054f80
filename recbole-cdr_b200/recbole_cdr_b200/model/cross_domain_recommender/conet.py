"""CoNet on the xdr hot path -- drop-in for reference model/cross_domain_recommender/conet.py.

Per tower pass: 2 concat-gathers (4 table reads) -> L cross-stitch layers, each ONE kernel per tower computing
``relu(W x + b + m * (x_other H_l))`` (the reference does 2 GEMMs, a boolean-indexed in-place add and a ReLU per
tower per layer, conet.py:118-138) -> output unit fused with BCE.  Same parameters and ``state_dict`` keys
(``source_crossunit_linear.{l}``, ``target_crossunit_linear.{l}``, ``crossparas.{l}``, ``*_outputunit.0``)."""
import os

import torch
import torch.nn as nn

from ... import _lib, ops
from ...utils import InputType
from ..crossdomain_recommender import CrossDomainRecommender
from ..init import xavier_normal_initialization


class CoNet(CrossDomainRecommender):
    input_type = InputType.POINTWISE

    def __init__(self, config, dataset):
        super(CoNet, self).__init__(config, dataset)
        self.SOURCE_LABEL = dataset.source_domain_dataset.label_field
        self.TARGET_LABEL = dataset.target_domain_dataset.label_field

        assert self.overlapped_num_items == 1 or self.overlapped_num_users == 1, \
            "CoNet model only support user overlapped or item overlapped dataset! "
        if self.overlapped_num_users > 1:
            self.mode = 'overlap_users'
        elif self.overlapped_num_items > 1:
            self.mode = 'overlap_items'
        else:
            self.mode = 'non_overlap'

        self.latent_dim = config['embedding_size']
        self.full_sort_block_bytes = 256 << 20   # full_sort_predict: bound on the hidden rows materialised at once
        self.reg_weight = config['reg_weight']  # read but never applied by the reference either (conet.py:53,198-201)
        self.cross_layers = list(config["mlp_hidden_size"])

        # construction order == reference order (conet.py:57-86)
        self.source_user_embedding = nn.Embedding(self.total_num_users, self.latent_dim)
        self.target_user_embedding = nn.Embedding(self.total_num_users, self.latent_dim)
        self.source_item_embedding = nn.Embedding(self.total_num_items, self.latent_dim)
        self.target_item_embedding = nn.Embedding(self.total_num_items, self.latent_dim)

        dims = [2 * self.latent_dim] + self.cross_layers
        self.source_crossunit_linear, self.source_crossunit_act = self.cross_units(dims)
        self.source_outputunit = nn.Sequential(nn.Linear(self.cross_layers[-1], 1), nn.Sigmoid())
        self.target_crossunit_linear, self.target_crossunit_act = self.cross_units(dims)
        self.target_outputunit = nn.Sequential(nn.Linear(self.cross_layers[-1], 1), nn.Sigmoid())
        self.crossparas = self.cross_parameters(dims)

        self.apply(xavier_normal_initialization)
        # opt-in: one tensor-core kernel per tower pass (tc_conet.cu) instead of the composed dense-layer kernels
        self.use_fused_conet = bool(config['xdr_fused_conet']) if 'xdr_fused_conet' in config else False
        # calculate_loss runs the source-batch and the target-batch tower passes as one pass over the stacked rows
        # (``xdr_stack_passes: False`` / XDR_CONET_STACK=0 keeps two passes)
        self.stack_passes = bool(config['xdr_stack_passes']) if 'xdr_stack_passes' in config else \
            os.environ.get('XDR_CONET_STACK', '1') != '0'
        # engine of the cross-stitch layers in calculate_loss (ops.dense_engine): 1 = tcgen05 (bf16x3 products, fp32-faithful) for
        # every layer shape it takes, the fp32 tiles for the rest.  Measured on a B200 at BASELINE configs[2] (stacked pass,
        # 2 x 16384 rows): 1005 us per step on engine 1, 1060 on engine 2, 1132 on engine 0 (profiles/r2_conet_stacked.md)
        if 'xdr_dense_engine' in config:
            self.dense_engine = int(config['xdr_dense_engine'])
        else:
            self.dense_engine = int(os.environ.get('XDR_DENSE_ENGINE') or 1)

    def _fused_ok(self):
        return self.use_fused_conet and ops.conet_fused_supported([2 * self.latent_dim] + self.cross_layers, self.latent_dim)

    def _fused_tower_loss(self, user, item, label, want):
        out = self.source_outputunit[0] if want == 'source' else self.target_outputunit[0]
        if self.mode == 'overlap_users':
            mask_on_item, n_ov = False, self.overlapped_num_users
        else:
            mask_on_item, n_ov = True, self.overlapped_num_items
        tabs = (self.source_user_embedding.weight, self.source_item_embedding.weight, self.target_user_embedding.weight,
                self.target_item_embedding.weight)
        return ops.conet_tower_loss(0 if want == 'source' else 1, mask_on_item, n_ov, user, item, label, tabs, out.weight,
                                    out.bias, [m.weight for m in self.source_crossunit_linear],
                                    [m.bias for m in self.source_crossunit_linear],
                                    [m.weight for m in self.target_crossunit_linear],
                                    [m.bias for m in self.target_crossunit_linear], [m.weight for m in self.crossparas])

    @staticmethod
    def cross_units(dims):
        lin = [nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])]
        act = [nn.ReLU() for _ in lin]
        return nn.ModuleList(lin), nn.ModuleList(act)

    @staticmethod
    def cross_parameters(dims):
        return nn.ModuleList([nn.Linear(a, b, bias=False) for a, b in zip(dims[:-1], dims[1:])])

    def _mask(self, user, item):
        """The overlapped rows of a batch, as (ids, bound): ``ids < bound`` (conet.py:113-116)."""
        if self.mode == 'overlap_users':
            return user.contiguous(), self.overlapped_num_users
        return item.contiguous(), self.overlapped_num_items

    def _stack(self, user, item):
        """Gathers + every cross-stitch layer but the last (conet.py:105-138), both towers; returns (x_s, x_t, mask ids, bound)."""
        x_s = ops.GatherConcat.apply(self.source_user_embedding.weight, self.source_item_embedding.weight, user, item)
        x_t = ops.GatherConcat.apply(self.target_user_embedding.weight, self.target_item_embedding.weight, user, item)
        mask_ids, mask_lt = self._mask(user, item)
        for l in range(len(self.source_crossunit_linear) - 1):
            fs, ft, h = self.source_crossunit_linear[l], self.target_crossunit_linear[l], self.crossparas[l].weight
            x_s, x_t = ops.cross_pair(x_s, x_t, fs.weight, fs.bias, ft.weight, ft.bias, h, mask_ids, mask_lt, _lib.ACT_RELU,
                                      self.dense_engine)
        return x_s, x_t, mask_ids, mask_lt

    def _head(self, x_s, x_t, mask_ids, mask_lt, want):
        """Last cross-stitch layer of tower ``want`` only (the other tower's last layer feeds nothing) + its output unit."""
        h = self.crossparas[-1].weight
        if want == 'source':
            fc, out, x, x_other = self.source_crossunit_linear[-1], self.source_outputunit[0], x_s, x_t
        else:
            fc, out, x, x_other = self.target_crossunit_linear[-1], self.target_outputunit[0], x_t, x_s
        x = ops.dense(x, fc.weight, fc.bias, _lib.ACT_RELU, x_other, h, mask_ids, mask_lt, self.dense_engine)
        return ops.dense(x, out.weight, out.bias, _lib.ACT_NONE).reshape(-1)

    def _towers(self, user, item, want):
        """Both towers through the cross-stitch stack (conet.py:105-138); returns the logit [B] of tower ``want``."""
        return self._head(*self._stack(user, item), want)

    def _both_passes(self, s_user, s_item, t_user, t_item):
        """``source_forward(source batch)`` and ``target_forward(target batch)`` of ``calculate_loss`` (conet.py:194-195) in ONE
        pass: the two tower passes apply the same gathers and the same layers (same weights) to different rows up to the last
        layer, so the rows of the two batches are stacked -- half the launches, twice the rows per launch -- and split again
        where the passes differ (last layer + output unit of the wanted tower).  Same sums, row for row."""
        s_user, s_item, t_user, t_item = (t.reshape(-1) for t in (s_user, s_item, t_user, t_item))
        n_s, n_t = s_user.numel(), t_user.numel()
        x_s, x_t, mask_ids, mask_lt = self._stack(torch.cat([s_user, t_user]), torch.cat([s_item, t_item]))
        (xs_s, xs_t), (xt_s, xt_t) = x_s.split([n_s, n_t]), x_t.split([n_s, n_t])
        m_s, m_t = mask_ids.split([n_s, n_t])
        return self._head(xs_s, xt_s, m_s, mask_lt, 'source'), self._head(xs_t, xt_t, m_t, mask_lt, 'target')

    def source_forward(self, user, item):
        return torch.sigmoid(self._towers(user, item, 'source'))

    def target_forward(self, user, item):
        return torch.sigmoid(self._towers(user, item, 'target'))

    def touched_rows(self, interaction):
        """(table parameter, ids) pairs read by ``calculate_loss``: every tower pass gathers all four tables
        (for trainer.RowSparseOptimizer)."""
        users = [interaction[self.SOURCE_USER_ID], interaction[self.TARGET_USER_ID]]
        items = [interaction[self.SOURCE_ITEM_ID], interaction[self.TARGET_ITEM_ID]]
        rows = []
        for tab in (self.source_user_embedding.weight, self.target_user_embedding.weight):
            rows += [(tab, u) for u in users]
        for tab in (self.source_item_embedding.weight, self.target_item_embedding.weight):
            rows += [(tab, i) for i in items]
        return rows

    def calculate_loss(self, interaction):
        """BCE(source tower, source batch) + BCE(target tower, target batch) + sum_l ||H_l||_F (conet.py:183-203)."""
        if self._fused_ok():
            loss_s = self._fused_tower_loss(interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID],
                                            interaction[self.SOURCE_LABEL], 'source')
            loss_t = self._fused_tower_loss(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID],
                                            interaction[self.TARGET_LABEL], 'target')
        else:
            if self.stack_passes:
                logit_s, logit_t = self._both_passes(interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID],
                                                     interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID])
            else:
                logit_s = self._towers(interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID], 'source')
                logit_t = self._towers(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID], 'target')
            loss_s, _ = ops.bce_logit(logit_s, interaction[self.SOURCE_LABEL])
            loss_t, _ = ops.bce_logit(logit_t, interaction[self.TARGET_LABEL])
        hs = [para.weight for para in self.crossparas]
        if len(hs) <= ops.FrobSum.MAX_MATS:
            reg_loss = ops.frob_sum(hs)     # one launch each way (torch: 7 kernels per matrix)
        else:
            reg_loss = 0
            for h in hs:
                reg_loss = reg_loss + torch.norm(h)
        return loss_s + loss_t + reg_loss

    def predict(self, interaction):
        """Target tower without cross terms (conet.py:205-220); returns [B, 1] like the reference."""
        with torch.no_grad():
            x = ops.GatherConcat.apply(self.target_user_embedding.weight, self.target_item_embedding.weight,
                                       interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID])
            for fc in self.target_crossunit_linear:
                x = ops.dense(x, fc.weight, fc.bias, _lib.ACT_RELU)
            out = self.target_outputunit[0]
            return ops.dense(x, out.weight, out.bias, _lib.ACT_SIGMOID)

    def _full_sort_blocks(self, interaction):
        """Yields ``(first user position, scores [block, n_items])`` of the target tower over every (user, target item) pair
        (conet.py:222-242).  The reference runs one Python iteration per user over ``[E_u || E_i]`` rows.  Layer 0 is linear in
        the concatenation, ``W0 [E_u || E_i] + b0 = (W0[:, :D] E_u + b0) + W0[:, D:] E_i``, so the item half is computed ONCE
        for all items and the user half once per user; a pair then costs one add + ReLU and the narrow tail layers (7x fewer
        multiply-adds at CoNet.yaml's stack).  Users are processed in blocks bounded by the hidden rows' footprint; no host
        sync."""
        user = interaction[self.TARGET_USER_ID].reshape(-1)
        n_items, D = self.target_num_items, self.latent_dim
        fc0 = self.target_crossunit_linear[0]
        item_part = ops.dense(self.target_item_embedding.weight[:n_items], fc0.weight[:, D:].contiguous())
        user_part = ops.dense(ops.gather_rows_raw(self.target_user_embedding.weight, user),
                              fc0.weight[:, :D].contiguous(), fc0.bias)
        width = item_part.shape[1]
        block = max(1, self.full_sort_block_bytes // (4 * width * n_items))   # users per block of layer-0 outputs
        out_fc = self.target_outputunit[0]
        for s in range(0, user.numel(), block):
            h = torch.relu(user_part[s:s + block].unsqueeze(1) + item_part.unsqueeze(0)).reshape(-1, width)
            for fc in list(self.target_crossunit_linear)[1:]:
                h = ops.dense(h, fc.weight, fc.bias, _lib.ACT_RELU)
            yield s, ops.dense(h, out_fc.weight, out_fc.bias, _lib.ACT_SIGMOID).reshape(-1, n_items)

    def full_sort_predict(self, interaction):
        """conet.py:222-242: the dense [B, n_items] score matrix of the target tower (see ``_full_sort_blocks``)."""
        with torch.no_grad():
            return torch.cat([sc for _, sc in self._full_sort_blocks(interaction)], dim=0)

    def full_sort_topk(self, interaction, k, hist_ptr=None, hist_ids=None):
        """``full_sort_predict`` + recbole's full-sort masking (PAD column, per-user history CSR) + ``topk`` per user block: the
        [B, n_items] matrix is never held, only one block of it.  Returns (scores [B, k], item positions [B, k])."""
        with torch.no_grad():
            scores, ids = [], []
            for s, sc in self._full_sort_blocks(interaction):
                sc[:, 0] = -float('inf')
                if hist_ptr is not None:
                    lo, hi = hist_ptr[s], hist_ptr[s + sc.shape[0]]
                    counts = hist_ptr[s + 1:s + sc.shape[0] + 1] - hist_ptr[s:s + sc.shape[0]]
                    rows = torch.repeat_interleave(torch.arange(sc.shape[0], device=sc.device), counts)
                    sc[rows, hist_ids[lo:hi]] = -float('inf')
                top = torch.topk(sc, k, dim=1)
                scores.append(top.values)
                ids.append(top.indices)
            return torch.cat(scores, dim=0), torch.cat(ids, dim=0)
