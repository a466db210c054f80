"""DeepAPF on the xdr hot path -- drop-in for reference model/cross_domain_recommender/deepapf.py.

Per tower (deepapf.py:68-146): gather the shared, the domain-only and the other side's row; attention logits of
(shared * other) and (only * other) through a two-layer MLP; softmax over the two (the shared one masked to -1e31 for
non-overlapped ids); convex combination; Linear(D, 1) on (combined * other); sigmoid; BCE.  Loss = source + target
(deepapf.py:158-175).  Gathers and their gradient scatter-adds are the xdr row kernels, the MLP layers the xdr dense
kernels, the output unit + BCE the fused sigmoid-BCE kernel; the two-way softmax and the products are element-wise torch
ops on [B, D] tensors.  Same parameters and ``state_dict`` keys (``user_mlp.0/2``, ``item_mlp.0/2``, ``predict_layer``)."""
import torch
import torch.nn as nn

from ... import _lib, ops
from ...utils import InputType
from ..crossdomain_recommender import CrossDomainRecommender
from ..init import xavier_normal_initialization


class DeepAPF(CrossDomainRecommender):
    input_type = InputType.POINTWISE

    def __init__(self, config, dataset):
        super(DeepAPF, self).__init__(config, dataset)
        self.SOURCE_LABEL = dataset.source_domain_dataset.label_field
        self.TARGET_LABEL = dataset.target_domain_dataset.label_field
        assert self.overlapped_num_items == 1 or self.overlapped_num_users == 1, \
            "DeepAPF model only support user overlapped or item overlapped dataset! "
        if self.overlapped_num_users > 1:
            self.mode = 'overlap_users'
        elif self.overlapped_num_items > 1:
            self.mode = 'overlap_items'
        else:
            self.mode = 'non_overlap'
        self.embedding_size = config['embedding_size']
        self.beta = config['beta'] if 'beta' in config else None  # read by the reference (deepapf.py:44), never used

        d = self.embedding_size
        self.source_user_embedding = nn.Embedding(self.total_num_users, d)
        self.target_user_embedding = nn.Embedding(self.total_num_users, d)
        self.share_user_embedding = nn.Embedding(self.total_num_users, d)
        self.source_item_embedding = nn.Embedding(self.total_num_items, d)
        self.target_item_embedding = nn.Embedding(self.total_num_items, d)
        self.share_item_embedding = nn.Embedding(self.total_num_items, d)
        # the reference also binds each MLP to ``self.seq`` (deepapf.py:56-62), which leaves ``seq`` as a second name of
        # ``item_mlp`` in the module tree: state_dict() carries ``seq.*`` next to ``item_mlp.*`` and named_parameters()
        # reports the item MLP under ``seq.*``.  Kept, so that checkpoints interchange.
        self.user_mlp = self.seq = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Linear(d, 1, bias=False))
        self.item_mlp = self.seq = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Linear(d, 1, bias=False))
        self.predict_layer = nn.Linear(d, 1, bias=False)
        self.apply(xavier_normal_initialization)

    def _attention(self, mlp, x):
        h = ops.dense(x, mlp[0].weight, mlp[0].bias, _lib.ACT_RELU)
        return ops.dense(h, mlp[2].weight, None, _lib.ACT_NONE)            # [B, 1]

    def _tower_logit(self, domain, user, item):
        if self.mode == 'overlap_users':
            share = ops.gather_rows(self.share_user_embedding.weight, user)
            only = ops.gather_rows(getattr(self, f'{domain}_user_embedding').weight, user)
            other = ops.gather_rows(getattr(self, f'{domain}_item_embedding').weight, item)
            mask = (user > self.overlapped_num_users).unsqueeze(-1)         # strict '>' as in deepapf.py:73
            mlp = self.user_mlp
        else:
            other = ops.gather_rows(getattr(self, f'{domain}_user_embedding').weight, user)
            share = ops.gather_rows(self.share_item_embedding.weight, item)
            only = ops.gather_rows(getattr(self, f'{domain}_item_embedding').weight, item)
            mask = (item > self.overlapped_num_items).unsqueeze(-1)
            mlp = self.item_mlp
        a_share = self._attention(mlp, share * other).masked_fill(mask, -1e31)
        a_only = self._attention(mlp, only * other)
        alpha = torch.softmax(torch.cat([a_share, a_only], dim=1), dim=1)   # [B, 2]
        combined = alpha[:, 0:1] * share + alpha[:, 1:2] * only
        return ops.dense((combined * other).contiguous(), self.predict_layer.weight, None, _lib.ACT_NONE).reshape(-1)

    def source_forward(self, user, item):
        return torch.sigmoid(self._tower_logit('source', user, item))

    def target_forward(self, user, item):
        return torch.sigmoid(self._tower_logit('target', user, item))

    def calculate_loss(self, interaction):
        logit_s = self._tower_logit('source', interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID])
        logit_t = self._tower_logit('target', interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID])
        loss_s, _ = ops.bce_logit(logit_s, interaction[self.SOURCE_LABEL])
        loss_t, _ = ops.bce_logit(logit_t, interaction[self.TARGET_LABEL])
        return loss_s + loss_t

    def touched_rows(self, interaction):
        su, si = interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID]
        tu, ti = interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID]
        rows = [(self.source_user_embedding.weight, su), (self.source_item_embedding.weight, si),
                (self.target_user_embedding.weight, tu), (self.target_item_embedding.weight, ti)]
        if self.mode == 'overlap_users':
            rows += [(self.share_user_embedding.weight, su), (self.share_user_embedding.weight, tu)]
        else:
            rows += [(self.share_item_embedding.weight, si), (self.share_item_embedding.weight, ti)]
        return rows

    def predict(self, interaction):
        with torch.no_grad():
            return self.target_forward(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID])
