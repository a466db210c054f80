"""CoNet on the xdr hot path -- drop-in for reference model/cross_domain_recommender/conet.py.

Per tower pass: 2 concat-gathers (4 table reads) -> L cross-stitch layers, each ONE kernel per tower computing
``relu(W x + b + m * (x_other H_l))`` (the reference does 2 GEMMs, a boolean-indexed in-place add and a ReLU per
tower per layer, conet.py:118-138) -> output unit fused with BCE.  Same parameters and ``state_dict`` keys
(``source_crossunit_linear.{l}``, ``target_crossunit_linear.{l}``, ``crossparas.{l}``, ``*_outputunit.0``)."""
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

    def _towers(self, user, item, want):
        """Both towers through the cross-stitch stack (conet.py:105-138); returns the logit [B] of tower ``want``."""
        x_s = ops.GatherConcat.apply(self.source_user_embedding.weight, self.source_item_embedding.weight, user, item)
        x_t = ops.GatherConcat.apply(self.target_user_embedding.weight, self.target_item_embedding.weight, user, item)
        if self.mode == 'overlap_users':
            mask_ids, mask_lt = user, self.overlapped_num_users
        else:
            mask_ids, mask_lt = item, self.overlapped_num_items
        mask_ids = mask_ids.contiguous()
        n_layers = len(self.source_crossunit_linear)
        for l in range(n_layers):
            fs, ft, h = self.source_crossunit_linear[l], self.target_crossunit_linear[l], self.crossparas[l].weight
            last = l == n_layers - 1
            h_s = h_t = None
            if not last or want == 'source':
                h_s = ops.dense(x_s, fs.weight, fs.bias, _lib.ACT_RELU, x_t, h, mask_ids, mask_lt)
            if not last or want == 'target':
                h_t = ops.dense(x_t, ft.weight, ft.bias, _lib.ACT_RELU, x_s, h, mask_ids, mask_lt)
            x_s, x_t = h_s, h_t
        if want == 'source':
            out = self.source_outputunit[0]
            return ops.dense(x_s, out.weight, out.bias, _lib.ACT_NONE).reshape(-1)
        out = self.target_outputunit[0]
        return ops.dense(x_t, out.weight, out.bias, _lib.ACT_NONE).reshape(-1)

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
            logit_s = self._towers(interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID], 'source')
            logit_t = self._towers(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID], 'target')
            loss_s, _ = ops.bce_logit(logit_s, interaction[self.SOURCE_LABEL])
            loss_t, _ = ops.bce_logit(logit_t, interaction[self.TARGET_LABEL])
        reg_loss = 0
        for para in self.crossparas:
            reg_loss = reg_loss + torch.norm(para.weight)
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
