"""SSCDR on the xdr hot path -- drop-in for reference model/cross_domain_recommender/sscdr.py.

SOURCE / TARGET phases: triplet-margin loss on length-clipped rows of (user, item+, item-) (sscdr.py:134-160); OVERLAP
phase: MSE(mapping(Es[idx]), Et[idx]) + lambda * triplet(Et[idx], mapping(Es[pos]), mapping(Es[neg])) with pos / neg drawn
on the host from the source interactions (sscdr.py:162-187).  Row gathers and their gradient scatter-adds are the xdr row
kernels, the tanh mapping MLP the xdr dense kernels, the map-loss tail the fused MSE-vs-gathered-rows kernel; the clipping,
distances and hinge are element-wise torch ops on [B, D] tensors.  Same parameters and ``state_dict`` keys."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...utils import InputType
from ..crossdomain_recommender import CrossDomainRecommender
from ..init import xavier_normal_initialization
from ..layers import MLPLayers


class SSCDR(CrossDomainRecommender):
    input_type = InputType.PAIRWISE

    def __init__(self, config, dataset):
        super(SSCDR, self).__init__(config, dataset)
        assert self.overlapped_num_items == 1 or self.overlapped_num_users == 1, \
            "SSCDR model only support user overlapped or item overlapped dataset! "
        if self.overlapped_num_users > 1:
            self.mode = 'overlap_users'
        elif self.overlapped_num_items > 1:
            self.mode = 'overlap_items'
        else:
            self.mode = 'non_overlap'
        self.phase = None
        self.embedding_size = config['embedding_size']
        self.lamda = config['lambda']
        self.margin = config['margin']
        self.mlp_hidden_size = list(config['mlp_hidden_size'])
        self.mapping_layer = MLPLayers(layers=[self.embedding_size] + self.mlp_hidden_size + [self.embedding_size],
                                       activation='tanh', dropout=0, bn=False)
        # interaction lists of the source domain, as CSR: row = user (overlap_users) or item (overlap_items)
        src = dataset.source_domain_dataset
        u = src.inter_feat[src.uid_field].numpy()
        i = src.inter_feat[src.iid_field].numpy()
        if self.mode == 'overlap_users':
            self._hist_ptr, self._hist_ids = self._csr(u, i, self.total_num_users)
            self._candidates = np.concatenate([np.arange(self.overlapped_num_items),
                                               np.arange(self.target_num_items, self.total_num_items)])
        elif self.mode == 'overlap_items':
            self._hist_ptr, self._hist_ids = self._csr(i, u, self.total_num_items)
            self._candidates = np.concatenate([np.arange(self.overlapped_num_users),
                                               np.arange(self.target_num_users, self.total_num_users)])

        self.source_user_embedding = nn.Embedding(self.total_num_users, self.embedding_size)
        self.source_item_embedding = nn.Embedding(self.total_num_items, self.embedding_size)
        self.target_user_embedding = nn.Embedding(self.total_num_users, self.embedding_size)
        self.target_item_embedding = nn.Embedding(self.total_num_items, self.embedding_size)
        # the reference's zero-fill of dead rows (sscdr.py:62-67) is overwritten by the init below (sscdr.py:73)
        self.apply(xavier_normal_initialization)

    @staticmethod
    def _csr(rows, cols, n_rows):
        """Per-row interacted ids in interaction order (the per-row lists of sscdr.py:75-90)."""
        order = np.argsort(rows, kind='stable')
        ptr = np.zeros(n_rows + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=n_rows), out=ptr[1:])
        return ptr, cols[order].astype(np.int64)

    def sample(self, ids, mode='user'):
        """One interacted and one non-interacted id per entry, drawn with NumPy's global RNG in the reference's call
        order (sscdr.py:92-122: per entry the negative first -- redrawn while it is in the history -- then the positive), so
        that the same ``np.random.seed`` yields the same draws.  Entries without history use [0] (PAD) as in the reference."""
        ids = ids.cpu().numpy()
        pos, neg = np.zeros_like(ids), np.zeros_like(ids)
        for n, key in enumerate(ids):
            hist = self._hist_ids[self._hist_ptr[key]:self._hist_ptr[key + 1]]
            if hist.size == 0:
                hist = np.zeros(1, dtype=np.int64)
            cand = np.random.choice(self._candidates, size=1)[0]
            while cand in hist:
                cand = np.random.choice(self._candidates, size=1)[0]
            pos[n] = np.random.choice(hist, size=1)[0]
            neg[n] = cand
        dev = self.source_user_embedding.weight.device
        return torch.from_numpy(pos).to(dev), torch.from_numpy(neg).to(dev)

    @staticmethod
    def embedding_normalize(e):
        """sscdr.py:124-129: rows whose SQUARED length exceeds 1 are divided by that squared length."""
        sq = torch.sum(e * e, dim=1, keepdim=True)
        return e / torch.where(sq > 1, sq, torch.ones_like(sq))

    @staticmethod
    def embedding_distance(a, b):
        return torch.sum((a - b) ** 2, dim=1)

    def _triplet(self, anchor, pos, neg):
        n = self.embedding_normalize
        return F.triplet_margin_loss(n(anchor), n(pos), n(neg), margin=self.margin)

    def set_phase(self, phase):
        self.phase = phase

    def _rec_loss(self, interaction, domain):
        ut = getattr(self, f'{domain}_user_embedding').weight
        it = getattr(self, f'{domain}_item_embedding').weight
        tag = domain.upper()
        return self._triplet(ops.gather_rows(ut, interaction[getattr(self, f'{tag}_USER_ID')]),
                             ops.gather_rows(it, interaction[getattr(self, f'{tag}_ITEM_ID')]),
                             ops.gather_rows(it, interaction[getattr(self, f'{tag}_NEG_ITEM_ID')]))

    def calculate_source_loss(self, interaction):
        return self._rec_loss(interaction, 'source')

    def calculate_target_loss(self, interaction):
        return self._rec_loss(interaction, 'target')

    def calculate_map_loss(self, interaction):
        idx = interaction[self.OVERLAP_ID].squeeze(1)
        if self.mode == 'overlap_users':
            src, tgt, other = self.source_user_embedding.weight, self.target_user_embedding.weight, self.source_item_embedding.weight
            pos, neg = self.sample(idx, mode='user')
        else:
            src, tgt, other = self.source_item_embedding.weight, self.target_item_embedding.weight, self.source_user_embedding.weight
            pos, neg = self.sample(idx, mode='item')
        loss_s = ops.mse_rows(self.mapping_layer(ops.gather_rows(src, idx)), tgt, idx)   # target rows are NOT detached
        loss_u = self._triplet(ops.gather_rows(tgt, idx), self.mapping_layer(ops.gather_rows(other, pos)),
                               self.mapping_layer(ops.gather_rows(other, neg)))
        return loss_s + self.lamda * loss_u

    def calculate_loss(self, interaction):
        if self.phase == 'SOURCE':
            return self.calculate_source_loss(interaction)
        if self.phase == 'OVERLAP':
            return self.calculate_map_loss(interaction)
        return self.calculate_target_loss(interaction)

    def touched_rows(self, interaction):
        if self.phase == 'OVERLAP':
            raise NotImplementedError('the OVERLAP phase of SSCDR draws its item rows inside calculate_loss')
        domain = 'source' if self.phase == 'SOURCE' else 'target'
        tag = domain.upper()
        ut = getattr(self, f'{domain}_user_embedding').weight
        it = getattr(self, f'{domain}_item_embedding').weight
        return [(ut, interaction[getattr(self, f'{tag}_USER_ID')]), (it, interaction[getattr(self, f'{tag}_ITEM_ID')]),
                (it, interaction[getattr(self, f'{tag}_NEG_ITEM_ID')])]

    # ---- inference -------------------------------------------------------------------------------------------
    def _overlap_phase_embeddings(self, user, item=None):
        """(user vectors, item vectors or None) of the OVERLAP/BOTH phase (sscdr.py:203-217)."""
        if self.mode == 'overlap_users':
            mapped = self.mapping_layer(ops.gather_rows_raw(self.source_user_embedding.weight, user))
            user_e = torch.where((user < self.overlapped_num_users).unsqueeze(1), mapped,
                                 ops.gather_rows_raw(self.target_user_embedding.weight, user))
            item_e = None if item is None else ops.gather_rows_raw(self.target_item_embedding.weight, item)
        else:
            user_e = ops.gather_rows_raw(self.target_user_embedding.weight, user)
            item_e = None
            if item is not None:
                mapped = self.mapping_layer(ops.gather_rows_raw(self.source_item_embedding.weight, item))
                item_e = torch.where((item < self.overlapped_num_items).unsqueeze(1), mapped,
                                     ops.gather_rows_raw(self.target_item_embedding.weight, item))
        return user_e, item_e

    def predict(self, interaction):
        with torch.no_grad():
            n = self.embedding_normalize
            if self.phase in ('SOURCE', 'TARGET'):
                domain = self.phase.lower()
                tag = self.phase
                user_e = ops.gather_rows_raw(getattr(self, f'{domain}_user_embedding').weight,
                                             interaction[getattr(self, f'{tag}_USER_ID')])
                item_e = ops.gather_rows_raw(getattr(self, f'{domain}_item_embedding').weight,
                                             interaction[getattr(self, f'{tag}_ITEM_ID')])
            else:
                user_e, item_e = self._overlap_phase_embeddings(interaction[self.TARGET_USER_ID],
                                                                interaction[self.TARGET_ITEM_ID])
            return -self.embedding_distance(n(user_e), n(item_e))

    def _full_sort_operands(self, interaction):
        """(normalised user-side vectors [B, D], normalised candidate item rows [n, D]) of full_sort_predict in the current
        phase (sscdr.py:222-252)."""
        n = self.embedding_normalize
        if self.phase == 'SOURCE':
            user_e = n(ops.gather_rows_raw(self.source_user_embedding.weight, interaction[self.SOURCE_USER_ID]))
            w = self.source_item_embedding.weight
            return user_e, torch.cat([n(w[:self.overlapped_num_items]), n(w[self.target_num_items:])], dim=0)
        if self.phase == 'TARGET':
            user_e = n(ops.gather_rows_raw(self.target_user_embedding.weight, interaction[self.TARGET_USER_ID]))
            return user_e, n(self.target_item_embedding.weight[:self.target_num_items])
        user_e, _ = self._overlap_phase_embeddings(interaction[self.TARGET_USER_ID])
        if self.mode == 'overlap_users':
            all_item_e = self.target_item_embedding.weight[:self.target_num_items]
        else:
            ov = self.mapping_layer(self.source_item_embedding.weight[:self.overlapped_num_items].contiguous())
            all_item_e = torch.cat([ov, self.target_item_embedding.weight[self.overlapped_num_items:self.target_num_items]], dim=0)
        return n(user_e), n(all_item_e)

    def full_sort_topk(self, interaction, k, hist_ptr=None, hist_ids=None, engine='mma'):
        """Fused ``full_sort_predict`` + PAD/history masking + ``topk``.  The score is a negative
        squared distance, ``-|u - i|^2 = 2 u.i - |i|^2 - |u|^2``: a dot product of the augmented rows ``[2u, -1, 0..]`` and
        ``[i, |i|^2, 0..]`` (width D + 8) minus a per-user constant, so the dot-product scoring kernel applies unchanged."""
        with torch.no_grad():
            user_e, all_item_e = self._full_sort_operands(interaction)
            B, D = user_e.shape
            ua = torch.zeros((B, D + 8), dtype=torch.float32, device=user_e.device)
            ua[:, :D] = 2 * user_e
            ua[:, D] = -1.0
            ia = torch.zeros((all_item_e.shape[0], D + 8), dtype=torch.float32, device=user_e.device)
            ia[:, :D] = all_item_e
            ia[:, D] = torch.sum(all_item_e ** 2, -1)
            sc, pos = ops.full_sort_topk(ua, ia, k, first_item=1, hist_ptr=hist_ptr, hist_ids=hist_ids, engine=engine)
            return sc - torch.sum(user_e ** 2, -1).view(-1, 1), pos

    def full_sort_predict(self, interaction):
        """sscdr.py:222-259: negative squared distances to every candidate item (library matmul, drop-in shape)."""
        with torch.no_grad():
            user_e, all_item_e = self._full_sort_operands(interaction)
            dist = -2 * torch.matmul(user_e, all_item_e.permute(1, 0))
            dist += torch.sum(user_e ** 2, -1).view(-1, 1)
            dist += torch.sum(all_item_e ** 2, -1).view(1, -1)
            return -dist.view(-1)
