"""DTCDR (NeuMF base model) on the xdr hot path -- drop-in for reference model/cross_domain_recommender/dtcdr.py.

``u = max(Es_u[u], Et_u[u])``, ``i = max(Es_i[i], Et_i[i])`` is one fused dual-table gather per side written straight
into the concatenated MLP input (the reference: 4 gathers, 2 maximum, 1 cat -- dtcdr.py:113-121); each MLP layer is one
xdr dense kernel; the sigmoid output unit is fused with BCE.  ``state_dict`` keys match the reference
(``{source,target}_mlp_layers.mlp_layers.{1,4}``, ``{source,target}_predict_layer``).

The DMF base model (dtcdr.py:69-101,127-175) builds dense multi-hot ``[B, n_items]`` matrices and is not a gather
path; it is out of scope and raises NotImplementedError.
"""
import torch
import torch.nn as nn

from ... import _lib, ops
from ...utils import InputType
from ..crossdomain_recommender import CrossDomainRecommender
from ..init import xavier_normal_initialization
from ..layers import MLPLayers


class DTCDR(CrossDomainRecommender):
    input_type = InputType.POINTWISE

    def __init__(self, config, dataset):
        super(DTCDR, self).__init__(config, dataset)
        self.SOURCE_LABEL = dataset.source_domain_dataset.label_field
        self.TARGET_LABEL = dataset.target_domain_dataset.label_field

        self.embedding_size = config['embedding_size']
        self.mlp_hidden_size = list(config['mlp_hidden_size'])
        self.dropout_prob = config['dropout_prob']
        self.base_model = config['base_model']
        self.alpha = config['alpha']
        assert self.base_model in ['NeuMF', 'DMF'], "based model {} is not supported! ".format(self.base_model)
        if self.base_model != 'NeuMF':
            raise NotImplementedError('DTCDR base_model DMF is outside the xdr hot-path scope (dense multi-hot matmul)')

        self.source_user_embedding = nn.Embedding(self.total_num_users, self.embedding_size)
        self.source_item_embedding = nn.Embedding(self.total_num_items, self.embedding_size)
        self.target_user_embedding = nn.Embedding(self.total_num_users, self.embedding_size)
        self.target_item_embedding = nn.Embedding(self.total_num_items, self.embedding_size)
        # the reference's -inf fill of dead rows (dtcdr.py:54-59) is overwritten by the init below (dtcdr.py:104)
        self.source_mlp_layers = MLPLayers([2 * self.embedding_size] + self.mlp_hidden_size, self.dropout_prob)
        self.source_predict_layer = nn.Linear(self.mlp_hidden_size[-1], 1)
        self.target_mlp_layers = MLPLayers([2 * self.embedding_size] + self.mlp_hidden_size, self.dropout_prob)
        self.target_predict_layer = nn.Linear(self.mlp_hidden_size[-1], 1)

        self.apply(xavier_normal_initialization)
        # False: composed kernels; True / 'fma': fp32 row-tile kernel; 'tc': tensor-core row-tile kernel
        self.fused_mlp_engine = ops.fused_mlp_engine(config['xdr_fused_mlp'] if 'xdr_fused_mlp' in config else False)
        self.use_fused_mlp = self.fused_mlp_engine is not None

    def _tower(self, domain):
        mlp = self.source_mlp_layers if domain == 'source' else self.target_mlp_layers
        out = self.source_predict_layer if domain == 'source' else self.target_predict_layer
        lins = [m for m in mlp.mlp_layers if isinstance(m, nn.Linear)] + [out]
        return [l.weight for l in lins], [l.bias for l in lins]

    def _fused_ok(self):
        dims = [2 * self.embedding_size] + self.mlp_hidden_size + [1]
        no_dropout = self.dropout_prob == 0 or not self.training
        return self.use_fused_mlp and no_dropout and ops.fused_mlp_supported(dims, self.fused_mlp_engine)

    def _tables(self):
        return (self.source_user_embedding.weight, self.target_user_embedding.weight, self.source_item_embedding.weight,
                self.target_item_embedding.weight, None)

    def _logit(self, user, item, domain):
        x = ops.GatherMax2Concat.apply(self.source_user_embedding.weight, self.target_user_embedding.weight,
                                       self.source_item_embedding.weight, self.target_item_embedding.weight, user, item)
        if domain == 'source':
            h, out = self.source_mlp_layers(x), self.source_predict_layer
        else:
            h, out = self.target_mlp_layers(x), self.target_predict_layer
        return ops.dense(h, out.weight, out.bias, _lib.ACT_NONE).reshape(-1)

    def neumf_forward(self, user, item, domain='source'):
        return torch.sigmoid(self._logit(user, item, domain))

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
        """alpha*BCE_s + (1-alpha)*BCE_t (dtcdr.py:177-191)."""
        if self._fused_ok():
            # per domain: ONE kernel forward and ONE backward (gather + max-combine + MLP + BCE + scatter)
            ws, bs = self._tower('source')
            loss_s = ops.fused_mlp_loss(1, 1, _lib.ACT_RELU, interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID],
                                        interaction[self.SOURCE_LABEL], self._tables(), ws, bs, self.fused_mlp_engine)
            wt, bt = self._tower('target')
            loss_t = ops.fused_mlp_loss(1, 1, _lib.ACT_RELU, interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID],
                                        interaction[self.TARGET_LABEL], self._tables(), wt, bt, self.fused_mlp_engine)
            return loss_s * self.alpha + loss_t * (1 - self.alpha)
        logit_s = self._logit(interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID], 'source')
        logit_t = self._logit(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID], 'target')
        loss_s, _ = ops.bce_logit(logit_s, interaction[self.SOURCE_LABEL])
        loss_t, _ = ops.bce_logit(logit_t, interaction[self.TARGET_LABEL])
        return loss_s * self.alpha + loss_t * (1 - self.alpha)

    def predict(self, interaction):
        with torch.no_grad():
            if self._fused_ok():
                wt, bt = self._tower('target')
                return ops.fused_mlp_prob(_lib.ACT_RELU, interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID],
                                          self._tables(), wt, bt, self.fused_mlp_engine)
            return self.neumf_forward(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID], 'target')
