"""CMF on the xdr hot path -- drop-in for reference model/cross_domain_recommender/cmf.py.

One shared user table and one shared item table; ``alpha*(BCE_s + lambda*Emb_s) + (1-alpha)*(BCE_t + gamma*Emb_t)``
(cmf.py:81-99).  Each domain term is ONE fused kernel forward (gather -> dot -> sigmoid -> BCE -> EmbLoss partials) and
one backward (re-gather -> row gradients -> vector-atomic scatter-add); the reference's per-domain term is 4 gathers,
a sigmoid, a BCE and two norms, plus dense [N, D] gradient tensors.
"""
import torch
import torch.nn as nn

from ... import _lib, ops
from ...utils import InputType
from ..crossdomain_recommender import CrossDomainRecommender
from ..init import xavier_normal_initialization


class CMF(CrossDomainRecommender):
    input_type = InputType.POINTWISE

    def __init__(self, config, dataset):
        super(CMF, self).__init__(config, dataset)
        self.SOURCE_LABEL = dataset.source_domain_dataset.label_field
        self.TARGET_LABEL = dataset.target_domain_dataset.label_field

        self.embedding_size = config['embedding_size']
        self.alpha = config['alpha']
        self.lamda = config['lambda']
        self.gamma = config['gamma']

        self.user_embedding = nn.Embedding(self.total_num_users, self.embedding_size)
        self.item_embedding = nn.Embedding(self.total_num_items, self.embedding_size)
        self.apply(xavier_normal_initialization)

    def get_user_embedding(self, user):
        return ops.gather_rows(self.user_embedding.weight, user)

    def get_item_embedding(self, item):
        return ops.gather_rows(self.item_embedding.weight, item)

    def forward(self, user, item):
        """sigmoid(dot) (cmf.py:75-79), inference form."""
        return torch.sigmoid(ops.dot_score(self.user_embedding.weight, self.item_embedding.weight, user, item))

    def calculate_loss(self, interaction):
        ut, it = self.user_embedding.weight, self.item_embedding.weight
        loss_s = ops.point_loss(ut, it, interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID],
                                interaction[self.SOURCE_LABEL], _lib.LOSS_BCE_SIGMOID, self.lamda)
        loss_t = ops.point_loss(ut, it, interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID],
                                interaction[self.TARGET_LABEL], _lib.LOSS_BCE_SIGMOID, self.gamma)
        return loss_s * self.alpha + loss_t * (1 - self.alpha)

    def touched_rows(self, interaction):
        """(table parameter, ids) pairs read by ``calculate_loss`` (for trainer.RowSparseOptimizer)."""
        ut, it = self.user_embedding.weight, self.item_embedding.weight
        return [(ut, interaction[self.SOURCE_USER_ID]), (ut, interaction[self.TARGET_USER_ID]),
                (it, interaction[self.SOURCE_ITEM_ID]), (it, interaction[self.TARGET_ITEM_ID])]

    def fused_step_spec(self):
        """Two weighted domain terms on the shared tables (cmf.py:81-99): the trainer's persistent multi-step path runs
        them as two launches per chunk, with loss weights alpha and 1 - alpha folded into the SGD scale."""
        ut, it = self.user_embedding.weight, self.item_embedding.weight
        return [dict(user_tab=ut, item_tab=it, pairwise=False, loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=self.lamda,
                     fields=[self.SOURCE_USER_ID, self.SOURCE_ITEM_ID], label_field=self.SOURCE_LABEL, loss_weight=self.alpha),
                dict(user_tab=ut, item_tab=it, pairwise=False, loss_kind=_lib.LOSS_BCE_SIGMOID, reg_weight=self.gamma,
                     fields=[self.TARGET_USER_ID, self.TARGET_ITEM_ID], label_field=self.TARGET_LABEL,
                     loss_weight=1 - self.alpha)]

    def predict(self, interaction):
        return self.forward(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID])

    def full_sort_predict(self, interaction):
        """cmf.py:107-112 (dense scoring GEMM: outside the training hot path, library matmul)."""
        with torch.no_grad():
            user_e = ops.gather_rows_raw(self.user_embedding.weight, interaction[self.TARGET_USER_ID])
            all_item_e = self.item_embedding.weight[:self.target_num_items]
            return torch.matmul(user_e, all_item_e.transpose(0, 1)).view(-1)

    def full_sort_topk(self, interaction, k, hist_ptr=None, hist_ids=None, engine='mma'):
        """Fused ``full_sort_predict`` + PAD/history masking + ``topk``: scores are the raw dot
        products of cmf.py:107-112 (no sigmoid there either); the [B, n_items] matrix is never written."""
        user_e = ops.gather_rows_raw(self.user_embedding.weight, interaction[self.TARGET_USER_ID])
        return ops.full_sort_topk(user_e, self.item_embedding.weight, k, n_items=self.target_num_items, first_item=1,
                                  hist_ptr=hist_ptr, hist_ids=hist_ids, engine=engine)
