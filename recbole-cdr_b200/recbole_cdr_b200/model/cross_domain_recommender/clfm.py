"""CLFM on the xdr hot path -- drop-in for reference model/cross_domain_recommender/clfm.py.

Per domain: user row -> [shared | domain-only] linear factors (ONE dense kernel on the stacked weights) -> dot with the item
row -> sigmoid -> BCE, plus reg_weight * EmbLoss(user rows, item rows); loss = alpha * source + (1 - alpha) * target
(clfm.py:70-122).  The dot + BCE (+ its backward and the item-gradient scatter-add) is the fused pair-score kernel with the
factor matrix standing in for the user table (row b of the batch = "user" b); the EmbLoss term is the same kernel's
regulariser-only form on the raw tables.  Same parameters and ``state_dict`` keys as the reference."""
import torch
import torch.nn as nn

from ... import _lib, ops
from ...utils import InputType
from ..crossdomain_recommender import CrossDomainRecommender
from ..init import xavier_normal_initialization


class CLFM(CrossDomainRecommender):
    input_type = InputType.POINTWISE

    def __init__(self, config, dataset):
        super(CLFM, self).__init__(config, dataset)
        self.SOURCE_LABEL = dataset.source_domain_dataset.label_field
        self.TARGET_LABEL = dataset.target_domain_dataset.label_field

        du, di, shared = config['user_embedding_size'], config['source_item_embedding_size'], config['share_embedding_size']
        # the reference reads 'source_item_embedding_size' for the target side as well (clfm.py:37): kept, so that the
        # same config builds the same shapes
        self.user_embedding_size, self.share_embedding_size = du, shared
        self.source_item_embedding_size = self.target_item_embedding_size = di
        self.alpha, self.reg_weight = config['alpha'], config['reg_weight']
        if not 0 <= shared <= di:
            raise AssertionError('share_embedding_size must lie in [0, item embedding size]')
        # registration order == the reference's (clfm.py:47-63): user tables, item tables, shared / source-only / target-only
        for side, rows, width in (('user', self.total_num_users, du), ('item', self.total_num_items, di)):
            for domain in ('source', 'target'):
                setattr(self, f'{domain}_{side}_embedding', nn.Embedding(rows, width))
        if shared > 0:
            self.shared_linear = nn.Linear(du, shared, bias=False)
        for domain in ('source', 'target'):
            if di - shared > 0:
                setattr(self, f'{domain}_only_linear', nn.Linear(du, di - shared, bias=False))
        self.apply(xavier_normal_initialization)

    def _factor_weight(self, domain):
        """[shared | only] stacked along the output dimension == torch.cat(factors, dim=1) of clfm.py:74-82."""
        parts = []
        if self.share_embedding_size > 0:
            parts.append(self.shared_linear.weight)
        only = getattr(self, f'{domain}_only_linear', None)
        if only is not None:
            parts.append(only.weight)
        return parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)

    def _tables(self, domain):
        if domain == 'source':
            return self.source_user_embedding.weight, self.source_item_embedding.weight
        return self.target_user_embedding.weight, self.target_item_embedding.weight

    def _factors(self, domain, user):
        ut, _ = self._tables(domain)
        return ops.dense(ops.gather_rows(ut, user), self._factor_weight(domain), None, _lib.ACT_NONE)

    def _logit(self, domain, user, item):
        f = self._factors(domain, user)
        rows = torch.arange(f.shape[0], device=f.device, dtype=torch.int64)
        return ops.dot_score(f, self._tables(domain)[1], rows, item)

    def source_forward(self, user, item):
        return torch.sigmoid(self._logit('source', user, item))

    def target_forward(self, user, item):
        return torch.sigmoid(self._logit('target', user, item))

    def _domain_loss(self, domain, user, item, label):
        ut, it = self._tables(domain)
        f = self._factors(domain, user)
        rows = torch.arange(f.shape[0], device=f.device, dtype=torch.int64)
        data = ops.point_loss(f, it, rows, item, label, _lib.LOSS_BCE_SIGMOID, 0.0)
        if ut.shape[1] == it.shape[1]:
            reg = ops.point_loss(ut, it, user, item, None, _lib.LOSS_NONE, self.reg_weight)
        else:  # EmbLoss over rows of different widths: (||U_b||_F + ||I_b||_F) / B
            reg = self.reg_weight * (torch.norm(ops.gather_rows(ut, user)) + torch.norm(ops.gather_rows(it, item))) / user.numel()
        return data + reg

    def calculate_loss(self, interaction):
        loss_s = self._domain_loss('source', interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID],
                                   interaction[self.SOURCE_LABEL])
        loss_t = self._domain_loss('target', interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID],
                                   interaction[self.TARGET_LABEL])
        return loss_s * self.alpha + loss_t * (1 - self.alpha)

    def touched_rows(self, interaction):
        return [(self.source_user_embedding.weight, interaction[self.SOURCE_USER_ID]),
                (self.source_item_embedding.weight, interaction[self.SOURCE_ITEM_ID]),
                (self.target_user_embedding.weight, interaction[self.TARGET_USER_ID]),
                (self.target_item_embedding.weight, interaction[self.TARGET_ITEM_ID])]

    def predict(self, interaction):
        with torch.no_grad():
            return self.target_forward(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID])

    def full_sort_predict(self, interaction):
        """clfm.py:130-145: factors x all target item rows (library matmul, drop-in shape)."""
        with torch.no_grad():
            f = self._factors('target', interaction[self.TARGET_USER_ID])
            return torch.matmul(f, self.target_item_embedding.weight[:self.target_num_items].transpose(0, 1)).view(-1)

    def full_sort_topk(self, interaction, k, hist_ptr=None, hist_ids=None, engine='mma'):
        f = self._factors('target', interaction[self.TARGET_USER_ID]).detach()
        return ops.full_sort_topk(f, self.target_item_embedding.weight, k, n_items=self.target_num_items, first_item=1,
                                  hist_ptr=hist_ptr, hist_ids=hist_ids, engine=engine)
