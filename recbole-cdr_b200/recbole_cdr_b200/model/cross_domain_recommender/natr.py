"""NATR on the xdr hot path -- drop-in for reference model/cross_domain_recommender/natr.py.

Phase 1 (SOURCE): BCE(sigmoid(dot(Es_u[u], Es_i[i]))) -- ONE fused pair-score kernel (natr.py:103-115).
Phase 2 (TARGET): the target-side row, an attention-pooled summary of up to ``max_inter_length`` transferred source rows
from the entity's target-domain history, a two-way domain attention, dot, sigmoid, BCE + reg_weight * RegLoss over the
target tables and the three layers (natr.py:117-172).  History rows are gathered with the xdr row kernel ([B * H] ids),
the transfer layer is one xdr dense kernel over [B * H, D]; the attention arithmetic is element-wise torch on [B, H, D].
The source tables are frozen in phase 2 (natr.py:79-83).  Same parameters and ``state_dict`` keys."""
import torch
import torch.nn as nn

from ... import _lib, ops
from ...utils import InputType
from ..crossdomain_recommender import CrossDomainRecommender
from ..init import xavier_normal_initialization


class NATR(CrossDomainRecommender):
    input_type = InputType.POINTWISE

    def __init__(self, config, dataset):
        super(NATR, self).__init__(config, dataset)
        assert self.overlapped_num_items == 1 or self.overlapped_num_users == 1, \
            "NATR model only support user overlapped or item overlapped dataset! "
        if self.overlapped_num_users > 1:
            self.mode = 'overlap_users'
        elif self.overlapped_num_items > 1:
            self.mode = 'overlap_items'
        else:
            self.mode = 'non_overlap'
        self.phase = None
        self.source_embedding_size = config['source_embedding_size']
        self.target_embedding_size = config['target_embedding_size']
        self.reg_weight = config['reg_weight']
        self.max_inter_length = config['max_inter_length']
        self.SOURCE_LABEL = dataset.source_domain_dataset.label_field
        self.TARGET_LABEL = dataset.target_domain_dataset.label_field
        # target-domain histories, truncated to max_inter_length columns (natr.py:85-101): overlapped ITEMS are reached
        # through a user's item history, overlapped USERS through an item's user history
        if self.mode == 'overlap_users':
            hist, _, lens = dataset.history_user_matrix(domain='target')
        elif self.mode == 'overlap_items':
            hist, _, lens = dataset.history_item_matrix(domain='target')
        if self.mode != 'non_overlap':
            hist = hist[:, :self.max_inter_length].contiguous()
            dev = config['device']
            self.history_matrix = hist.to(dev)
            self.history_lens = lens.to(dev)
            self.mask_mat = (torch.arange(hist.shape[1]) < lens.unsqueeze(1)).float().to(dev)

        ds, dt = self.source_embedding_size, self.target_embedding_size
        # registration order == the reference's (natr.py:60-72): source user/item, target user/item, then the three layers
        for domain, width in (('source', ds), ('target', dt)):
            for side, rows in (('user', self.total_num_users), ('item', self.total_num_items)):
                setattr(self, f'{domain}_{side}_embedding', nn.Embedding(rows, width))
        self.transfer_layer = nn.Linear(ds, dt)
        for name in ('unit_attention_layer', 'domain_attention_layer'):
            setattr(self, name, nn.Linear(dt, 1))
        # the zero-fill of dead rows (natr.py:64-68) is overwritten by the init below (natr.py:77)
        self.apply(xavier_normal_initialization)

    def set_phase(self, phase):
        self.phase = phase
        if phase == 'TARGET':
            self.source_item_embedding.weight.requires_grad = False
            self.source_user_embedding.weight.requires_grad = False

    def _phase2_logit(self, user, item):
        user_e = ops.gather_rows(self.target_user_embedding.weight, user)
        item_e = ops.gather_rows(self.target_item_embedding.weight, item)
        if self.mode == 'overlap_items':
            key, src_tab, pu, qi = user, self.source_item_embedding.weight, user_e, item_e
        else:
            key, src_tab, pu, qi = item, self.source_user_embedding.weight, item_e, user_e
        bias = torch.where(self.mask_mat[key].bool(), 0., -10000.0)                       # [B, H]
        hist = self.history_matrix[key]                                                   # [B, H]
        B, H = hist.shape
        h = ops.dense(ops.gather_rows(src_tab, hist.reshape(-1)), self.transfer_layer.weight, self.transfer_layer.bias,
                      _lib.ACT_NONE).view(B, H, -1)                                       # transferred history rows
        att = torch.relu(pu.unsqueeze(1) * h)
        att = ops.dense(att.reshape(B * H, -1), self.unit_attention_layer.weight, self.unit_attention_layer.bias,
                        _lib.ACT_NONE).view(B, H) + bias
        su = torch.bmm(torch.softmax(att, dim=1).unsqueeze(1), h).squeeze(1)              # [B, D]
        dw, db = self.domain_attention_layer.weight, self.domain_attention_layer.bias
        b_s = ops.dense(torch.relu(su * qi), dw, db, _lib.ACT_NONE)
        b_p = ops.dense(torch.relu(pu * qi), dw, db, _lib.ACT_NONE)
        beta_s = torch.exp(b_s) / (torch.exp(b_s) + torch.exp(b_p))
        zu = beta_s * su + (1 - beta_s) * pu
        return (zu * qi).sum(dim=1)

    def phase1_forward(self, user, item):
        return torch.sigmoid(ops.dot_score(self.source_user_embedding.weight, self.source_item_embedding.weight, user, item))

    def phase2_forward(self, user, item):
        return torch.sigmoid(self._phase2_logit(user, item))

    def calculate_phase1_loss(self, interaction):
        return ops.point_loss(self.source_user_embedding.weight, self.source_item_embedding.weight,
                              interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID],
                              interaction[self.SOURCE_LABEL], _lib.LOSS_BCE_SIGMOID, 0.0).reshape(())

    def calculate_phase2_loss(self, interaction):
        logit = self._phase2_logit(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID])
        rec_loss, _ = ops.bce_logit(logit, interaction[self.TARGET_LABEL])
        reg = None
        for w in (self.target_user_embedding.weight, self.target_item_embedding.weight, self.transfer_layer.weight,
                  self.unit_attention_layer.weight, self.domain_attention_layer.weight):
            reg = w.norm(2) if reg is None else reg + w.norm(2)                           # recbole RegLoss
        return rec_loss + self.reg_weight * reg

    def calculate_loss(self, interaction):
        if self.phase == 'SOURCE':
            return self.calculate_phase1_loss(interaction)
        if self.phase == 'TARGET':
            return self.calculate_phase2_loss(interaction)
        return None

    def predict(self, interaction):
        with torch.no_grad():
            if self.phase == 'SOURCE':
                return self.phase1_forward(interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID])
            return self.phase2_forward(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID])
