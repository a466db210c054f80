"""DCDCSR on the xdr hot path -- drop-in for reference model/cross_domain_recommender/dcdcsr.py.

Four stages keyed by (phase, how often that phase has been entered) (dcdcsr.py:100-118, 216-240):
  SOURCE #1 / TARGET #1   BPR on the raw tables                      -> ONE fused BPR kernel (reg_weight 0)
  BOTH                    benchmark embedding per target unit (built once on entering the phase) and
                          MSE(mapping(minmax(Et[s])), minmax(benchmark[s])) on a random sample s of map_batch_size units
  TARGET #2               BPR with the affine (mapped, frozen) embedding on the overlapped side
The reference builds the benchmark with one Python iteration per unit (a [n_source, D] x [D] product and a top-k each,
dcdcsr.py:136-159); here the non-overlapped units are scored against the source rows in ONE pass of the fused
score + top-k kernel (``ops.full_sort_topk``) and combined with vector ops.  Same parameters and ``state_dict`` keys."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...utils import InputType
from ..crossdomain_recommender import CrossDomainRecommender
from ..init import xavier_normal_initialization
from ..layers import MLPLayers


class DCDCSR(CrossDomainRecommender):
    input_type = InputType.PAIRWISE

    def __init__(self, config, dataset):
        super(DCDCSR, self).__init__(config, dataset)
        assert self.overlapped_num_items == 1 or self.overlapped_num_users == 1, \
            "DCDCSR model only support user overlapped or item overlapped dataset! "
        if self.overlapped_num_users > 1:
            self.mode = 'overlap_users'
        elif self.overlapped_num_items > 1:
            self.mode = 'overlap_items'
        else:
            self.mode = 'non_overlap'
        self.phase = None
        self.phase2count = {'SOURCE': 0, 'TARGET': 0, 'BOTH': 0, 'OVERLAP': 0}

        self.latent_factor_model = config['latent_factor_model']
        assert self.latent_factor_model in ['BPR'], "latent_factor model must be in [BPR]"
        self.embedding_size = config['embedding_size']
        self.mlp_hidden_size = list(config['mlp_hidden_size'])
        self.k = config['k']
        self.map_batch_size = config['map_batch_size']
        self.bpr_gamma = 1e-10  # recbole BPRLoss default

        self.SOURCE_LABEL = dataset.source_domain_dataset.label_field
        self.TARGET_LABEL = dataset.target_domain_dataset.label_field
        # popularity (history length) of every unit on the overlapped side, per domain (dcdcsr.py:61-66, 92-98)
        if self.mode == 'overlap_items':
            self.source_unit2pop = dataset.history_user_matrix(domain='source')[2].float()
            self.target_unit2pop = dataset.history_user_matrix(domain='target')[2].float()
        elif self.mode == 'overlap_users':
            self.source_unit2pop = dataset.history_item_matrix(domain='source')[2].float()
            self.target_unit2pop = dataset.history_item_matrix(domain='target')[2].float()

        self.source_user_embedding = nn.Embedding(self.total_num_users, self.embedding_size)
        self.source_item_embedding = nn.Embedding(self.total_num_items, self.embedding_size)
        self.target_user_embedding = nn.Embedding(self.total_num_users, self.embedding_size)
        self.target_item_embedding = nn.Embedding(self.total_num_items, self.embedding_size)
        self.benchmark_embedding = None
        self.affine_embedding = None
        self.mapping_mlp_layers = MLPLayers(layers=[self.embedding_size] + self.mlp_hidden_size + [self.embedding_size],
                                            activation='tanh', dropout=0, bn=False)
        # the zero-fill of dead rows (dcdcsr.py:74-79) is overwritten by the init below (dcdcsr.py:89)
        self.apply(xavier_normal_initialization)

    # ---- the overlapped side ------------------------------------------------------------------------------------
    def _side(self):
        """(source table, target table, n overlapped, n target units, n total units) of the overlapped side."""
        if self.mode == 'overlap_users':
            return (self.source_user_embedding.weight, self.target_user_embedding.weight, self.overlapped_num_users,
                    self.target_num_users, self.total_num_users)
        return (self.source_item_embedding.weight, self.target_item_embedding.weight, self.overlapped_num_items,
                self.target_num_items, self.total_num_items)

    @staticmethod
    def maxmin_normalize(w):
        mn = torch.amin(w, dim=1, keepdim=True)
        mx = torch.amax(w, dim=1, keepdim=True)
        mean = (mx + mn) / 2
        return (w - mean) / (mx - mean), mean, mx

    def set_phase(self, phase):
        self.phase = phase
        self.phase2count[phase] += 1
        if phase == 'BOTH':
            self.build_benchmark_embedding()
        if phase == 'TARGET' and self.phase2count[phase] == 2:
            with torch.no_grad():
                _, tgt, _, n_tgt, _ = self._side()
                normal, mean, mx = self.maxmin_normalize(tgt[:n_tgt])
                self.affine_embedding = (self.mapping_mlp_layers(normal.contiguous()) * (mx - mean) + mean).detach()

    def build_benchmark_embedding(self):
        """dcdcsr.py:136-170.  Overlapped unit u: popularity-weighted mix of its two rows.  Any other unit i: its target row
        mixed with the similarity-weighted mean of its k most similar overlapped SOURCE rows (similarity = dot product),
        mixing weight beta = mean source popularity of those k / (that + target popularity of i)."""
        with torch.no_grad():
            src_tab, tgt_tab, n_ov, _, n_total = self._side()
            dev = tgt_tab.device
            src = src_tab[:n_ov].contiguous()
            pop_s, pop_t = self.source_unit2pop.to(dev), self.target_unit2pop.to(dev)
            bench = torch.empty((n_total, self.embedding_size), dtype=torch.float32, device=dev)
            den = pop_s[:n_ov] + pop_t[:n_ov]
            den = torch.where(den == 0, torch.ones_like(den), den)
            a_s = (pop_s[:n_ov] / den).unsqueeze(1)
            bench[:n_ov] = a_s * tgt_tab[:n_ov] + (1 - a_s) * src
            if n_total > n_ov:
                rest = tgt_tab[n_ov:].contiguous()
                sim, index = ops.full_sort_topk(rest, src, self.k, first_item=0)       # [n_rest, k]: no PAD skip here
                sn = pop_s[index].mean(dim=1)
                beta = (sn / (sn + pop_t[n_ov:])).unsqueeze(1)
                sim_e = torch.bmm(sim.unsqueeze(1), src[index]).squeeze(1)             # sum_k sim_k * row_k
                tot = sim.sum(dim=1, keepdim=True)
                sim_e = sim_e / torch.where(tot > 0, tot, torch.ones_like(tot))
                bench[n_ov:] = (1 - beta) * rest + beta * sim_e
            self.benchmark_embedding = bench

    # ---- losses ------------------------------------------------------------------------------------------------
    def _bpr(self, interaction, user_tab, item_tab, domain):
        tag = domain.upper()
        return ops.bpr_loss(user_tab, item_tab, interaction[getattr(self, f'{tag}_USER_ID')],
                            interaction[getattr(self, f'{tag}_ITEM_ID')], interaction[getattr(self, f'{tag}_NEG_ITEM_ID')],
                            0.0, self.bpr_gamma).reshape(())

    def calculate_map_loss(self):
        _, tgt_tab, _, n_tgt, _ = self._side()
        sampled = torch.from_numpy(np.random.randint(0, n_tgt, self.map_batch_size)).to(tgt_tab.device)  # dcdcsr.py:180
        rows, _, _ = self.maxmin_normalize(ops.gather_rows(tgt_tab, sampled))
        mapped = self.mapping_mlp_layers(rows.contiguous())
        bench, _, _ = self.maxmin_normalize(self.benchmark_embedding[sampled])
        return F.mse_loss(mapped, bench)

    def _stage(self):
        if self.phase == 'SOURCE' and self.phase2count['SOURCE'] == 1:
            return 'source1'
        if self.phase == 'TARGET' and self.phase2count['TARGET'] == 1:
            return 'target1'
        if self.phase == 'BOTH':
            return 'both'
        if self.phase == 'TARGET' and self.phase2count['TARGET'] == 2:
            return 'target2'
        return 'other'

    def _target2_tables(self):
        if self.mode == 'overlap_users':
            return self.affine_embedding, self.target_item_embedding.weight
        return self.target_user_embedding.weight, self.affine_embedding

    def calculate_loss(self, interaction):
        stage = self._stage()
        if stage == 'source1':
            return self._bpr(interaction, self.source_user_embedding.weight, self.source_item_embedding.weight, 'source')
        if stage == 'target1':
            return self._bpr(interaction, self.target_user_embedding.weight, self.target_item_embedding.weight, 'target')
        if stage == 'both':
            return self.calculate_map_loss()
        if stage == 'target2':
            ut, it = self._target2_tables()
            return self._bpr(interaction, ut, it, 'target')
        return None

    # ---- inference ------------------------------------------------------------------------------------------------
    def _eval_tables(self):
        """(user table, item table, user field, item field, all-item rows) of the current stage (dcdcsr.py:204-280)."""
        stage = self._stage()
        if stage == 'source1':
            w = self.source_item_embedding.weight
            return (self.source_user_embedding.weight, w, self.SOURCE_USER_ID, self.SOURCE_ITEM_ID,
                    lambda: torch.cat([w[:self.overlapped_num_items], w[self.target_num_items:]], dim=0))
        if stage == 'target1':
            w = self.target_item_embedding.weight
            return (self.target_user_embedding.weight, w, self.TARGET_USER_ID, self.TARGET_ITEM_ID,
                    lambda: w[:self.target_num_items])
        ut, it = self._target2_tables()
        if self.mode == 'overlap_users':
            return ut, it, self.TARGET_USER_ID, self.TARGET_ITEM_ID, lambda: it[:self.target_num_items]
        return ut, it, self.TARGET_USER_ID, self.TARGET_ITEM_ID, lambda: it

    def predict(self, interaction):
        with torch.no_grad():
            ut, it, uf, itf, _ = self._eval_tables()
            return ops.dot_score(ut.contiguous(), it.contiguous(), interaction[uf], interaction[itf])

    def full_sort_predict(self, interaction):
        """[B, n_items] (the reference does not flatten here, dcdcsr.py:241)."""
        with torch.no_grad():
            ut, _, uf, _, all_items = self._eval_tables()
            user_e = ops.gather_rows_raw(ut.contiguous(), interaction[uf])
            return torch.matmul(user_e, all_items().transpose(0, 1))

    def full_sort_topk(self, interaction, k, hist_ptr=None, hist_ids=None, engine='mma'):
        """Fused ``full_sort_predict`` + PAD/history masking + ``topk`` in the current stage: the
        [B, n_items] matrix is never written.  Returns (scores [B, k], positions [B, k] in full_sort_predict's columns)."""
        with torch.no_grad():
            ut, _, uf, _, all_items = self._eval_tables()
            user_e = ops.gather_rows_raw(ut.contiguous(), interaction[uf])
            return ops.full_sort_topk(user_e, all_items().contiguous(), k, first_item=1, hist_ptr=hist_ptr, hist_ids=hist_ids,
                                      engine=engine)
