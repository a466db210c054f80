"""BiTGCF on the xdr hot path -- drop-in for reference model/cross_domain_recommender/bitgcf.py.

Per training batch the reference propagates the FULL graph of both domains (``calculate_loss`` calls ``forward()``,
bitgcf.py:209): ``n_layers`` x 2 sparse matmuls with ``E + S + E*S`` epilogues, a transfer step that cuts and re-concatenates
~14 full-size tensors, two normalisations, a layer combine, then four gathers and the loss.  Here:

  graph_layer     -> xdr_spmm_csr over a work-item cut of the CSR + one element-wise kernel (graph.GraphProp)
  transfer + norm -> ONE kernel over both domains (graph.TransferNorm)
  gather + sigmoid(dot) + BCE and the EmbLoss on the ego rows -> the fused point-loss kernels (ops.point_loss)

Same parameters and ``state_dict`` keys as the reference; ``other_parameter_name`` as bitgcf.py:90.  Dropout
(``drop_rate``) is applied with torch's ``F.dropout`` when non-zero (parity tests use ``drop_rate = 0``).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _lib, ops
from ...graph import GraphProp, NormAdj, TransferNorm
from ...utils import InputType
from ..crossdomain_recommender import CrossDomainRecommender
from ..init import xavier_normal_initialization


class BiTGCF(CrossDomainRecommender):
    input_type = InputType.POINTWISE

    def __init__(self, config, dataset):
        super(BiTGCF, self).__init__(config, dataset)
        self.SOURCE_LABEL = dataset.source_domain_dataset.label_field
        self.TARGET_LABEL = dataset.target_domain_dataset.label_field

        self.latent_dim = config['embedding_size']
        self.n_layers = config['n_layers']
        self.reg_weight = config['reg_weight']
        self.domain_lambda_source = config['lambda_source']
        self.domain_lambda_target = config['lambda_target']
        self.drop_rate = config['drop_rate']
        self.connect_way = config['connect_way']

        # construction order == reference order (bitgcf.py:53-57)
        self.source_user_embedding = nn.Embedding(self.total_num_users, self.latent_dim)
        self.target_user_embedding = nn.Embedding(self.total_num_users, self.latent_dim)
        self.source_item_embedding = nn.Embedding(self.total_num_items, self.latent_dim)
        self.target_item_embedding = nn.Embedding(self.total_num_items, self.latent_dim)

        device = config['device']
        src = dataset.inter_matrix(form='coo', value_field=None, domain='source')
        tgt = dataset.inter_matrix(form='coo', value_field=None, domain='target')
        self.source_norm_adj = NormAdj(src.row, src.col, self.total_num_users, self.total_num_items, device)
        self.target_norm_adj = NormAdj(tgt.row, tgt.col, self.total_num_users, self.total_num_items, device)
        # per-node degree vectors [n_users + n_items] (bitgcf.py:79-82: inter_matrix.sum(axis=1) / sum(axis=0))
        def degrees(m):
            du = np.asarray(m.astype(np.float32).sum(axis=1)).reshape(-1)
            di = np.asarray(m.astype(np.float32).sum(axis=0)).reshape(-1)
            return torch.from_numpy(np.concatenate([du, di]).astype(np.float32)).to(device)
        self.source_degree = degrees(src)
        self.target_degree = degrees(tgt)

        self.target_restore_user_e = None
        self.target_restore_item_e = None

        self.apply(xavier_normal_initialization)
        self.other_parameter_name = ['target_restore_user_e', 'target_restore_item_e']

    def forward(self):
        """bitgcf.py:174-205: per layer propagate both domains -> transfer -> normalise; combine; split users/items."""
        es = torch.cat([self.source_user_embedding.weight, self.source_item_embedding.weight], dim=0)
        et = torch.cat([self.target_user_embedding.weight, self.target_item_embedding.weight], dim=0)
        ls, lt = [es], [et]
        for _ in range(self.n_layers):
            ps = GraphProp.apply(es, self.source_norm_adj)
            pt = GraphProp.apply(et, self.target_norm_adj)
            if self.training and self.drop_rate > 0:
                ps, pt = F.dropout(ps, self.drop_rate, True), F.dropout(pt, self.drop_rate, True)
            es, et, ns, nt = TransferNorm.apply(ps, pt, self.source_degree, self.target_degree, self.total_num_users,
                                                self.total_num_items, self.overlapped_num_users, self.overlapped_num_items,
                                                self.domain_lambda_source, self.domain_lambda_target)
            ls.append(ns)
            lt.append(nt)
        if self.connect_way == 'concat':
            fs, ft = torch.cat(ls, 1), torch.cat(lt, 1)
        elif self.connect_way == 'mean':
            # mean over the layer outputs (bitgcf.py:195-198: stack + mean) as a running sum: stack + mean writes an
            # [N, L + 1, D] copy forward and, backward, an expanded gradient plus one strided-to-contiguous copy per layer --
            # 26 table-sized passes per domain at 3 layers where the running sum needs 13 (same value up to fp32 rounding)
            fs, ft = self._layer_mean(ls), self._layer_mean(lt)
        else:
            raise ValueError(f'connect_way [{self.connect_way}] is not supported')
        su, si = torch.split(fs, [self.total_num_users, self.total_num_items])
        tu, ti = torch.split(ft, [self.total_num_users, self.total_num_items])
        return su, si, tu, ti

    @staticmethod
    def _layer_mean(layers):
        acc = layers[0]
        for e in layers[1:]:
            acc = acc + e
        return acc * (1.0 / len(layers))

    def calculate_loss(self, interaction):
        """bitgcf.py:207-250: per domain BCE(sigmoid(dot of propagated rows)) + reg_weight * EmbLoss(ego rows);
        returns the TUPLE (source_loss, target_loss), each of shape [1]."""
        self.init_restore_e()
        su, si, tu, ti = self.forward()
        losses = []
        for ua, ia, uid, iid, lab, ue, ie in (
                (su, si, self.SOURCE_USER_ID, self.SOURCE_ITEM_ID, self.SOURCE_LABEL, self.source_user_embedding.weight,
                 self.source_item_embedding.weight),
                (tu, ti, self.TARGET_USER_ID, self.TARGET_ITEM_ID, self.TARGET_LABEL, self.target_user_embedding.weight,
                 self.target_item_embedding.weight)):
            user, item, label = interaction[uid], interaction[iid], interaction[lab]
            bce = ops.point_loss(ua.contiguous(), ia.contiguous(), user, item, label, _lib.LOSS_BCE_SIGMOID, 0.0)
            reg = ops.point_loss(ue, ie, user, item, None, _lib.LOSS_NONE, self.reg_weight)
            losses.append(bce + reg)
        return tuple(losses)

    def predict(self, interaction):
        with torch.no_grad():
            _, _, tu, ti = self.forward()
            return ops.dot_score(tu.contiguous(), ti.contiguous(), interaction[self.TARGET_USER_ID],
                                 interaction[self.TARGET_ITEM_ID])

    def full_sort_predict(self, interaction):
        """bitgcf.py:264-272 (dense scoring GEMM: outside the training hot path, library matmul)."""
        with torch.no_grad():
            user = interaction[self.TARGET_USER_ID]
            restore_user_e, restore_item_e = self.get_restore_e()
            u = ops.gather_rows_raw(restore_user_e.contiguous(), user)
            return torch.matmul(u, restore_item_e[:self.target_num_items].transpose(0, 1)).view(-1)

    def full_sort_topk(self, interaction, k, hist_ptr=None, hist_ids=None, engine='mma'):
        """Fused ``full_sort_predict`` + PAD/history masking + ``topk`` on the propagated tables: one
        scoring kernel that never writes the [B, n_items] matrix.  Row width is D (mean) or D * (n_layers + 1) (concat);
        the fused kernel takes widths up to 256."""
        with torch.no_grad():
            restore_user_e, restore_item_e = self.get_restore_e()
            u = ops.gather_rows_raw(restore_user_e.contiguous(), interaction[self.TARGET_USER_ID])
            return ops.full_sort_topk(u, restore_item_e.contiguous(), k, n_items=self.target_num_items, first_item=1,
                                      hist_ptr=hist_ptr, hist_ids=hist_ids, engine=engine)

    def init_restore_e(self):
        if self.target_restore_user_e is not None or self.target_restore_item_e is not None:
            self.target_restore_user_e, self.target_restore_item_e = None, None

    def get_restore_e(self):
        if self.target_restore_user_e is None or self.target_restore_item_e is None:
            with torch.no_grad():
                _, _, self.target_restore_user_e, self.target_restore_item_e = self.forward()
        return self.target_restore_user_e, self.target_restore_item_e
