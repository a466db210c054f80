"""EMCDR on the xdr hot path -- drop-in for reference model/cross_domain_recommender/emcdr.py.

Same constructor signature, class/instance attributes, ``state_dict`` keys (``source_user_embedding.weight`` ...
``mapping.0.weight``) and method semantics; the per-batch arithmetic runs in libxdr:

  calculate_source_loss / calculate_target_loss (emcdr.py:110-154)  -> ops.bpr_loss / ops.point_loss  (1 fused kernel
        forward, 1 backward; the reference does 6 gathers + 6 [B, D] temporaries + 2 dense [N, D] grads)
  calculate_map_loss (emcdr.py:156-168)   -> gather -> xdr dense layers -> fused MSE-vs-gathered-target
  predict (emcdr.py:178-206)              -> fused dot score, or gather -> dense -> select+dot
"""
import torch
import torch.nn as nn

from ... import _lib, ops
from ...utils import InputType
from ..crossdomain_recommender import CrossDomainRecommender
from ..init import xavier_normal_initialization


class EMCDR(CrossDomainRecommender):
    r"""EMCDR learns a mapping function from the source latent space to the target latent space
    (Man et al., IJCAI 2017)."""

    def __init__(self, config, dataset):
        super(EMCDR, self).__init__(config, dataset)

        assert self.overlapped_num_items == 1 or self.overlapped_num_users == 1, \
            "EMCDR model only support user overlapped or item overlapped dataset! "
        if self.overlapped_num_users > 1:
            self.mode = 'overlap_users'
        elif self.overlapped_num_items > 1:
            self.mode = 'overlap_items'
        else:
            self.mode = 'non_overlap'
        self.phase = 'both'

        self.latent_factor_model = config['latent_factor_model']
        if self.latent_factor_model == 'MF':
            self.input_type = InputType.POINTWISE
            self.SOURCE_LABEL = dataset.source_domain_dataset.label_field
            self.TARGET_LABEL = dataset.target_domain_dataset.label_field
        else:
            self.input_type = InputType.PAIRWISE
        self.bpr_gamma = 1e-10  # recbole BPRLoss default
        # Engine of the map step (config key ``xdr_fused_mlp``).  Absent / 'auto' (default): the tcgen05 kernel (tc5_mlp.cu:
        # gather -> both layers -> MSE -> whole backward -> scatter in one launch, bf16x3 products on the tensor cores) for
        # the stacks it takes -- [D, 128, D], D a multiple of 16 up to 64, i.e. the yaml default -- and the composed fp32
        # kernels for everything else (measured on a B200 at b = 8192: 37 us against 98 us per step, profiles/r2_rows.md).
        # False: always composed; True / 'fma': fp32 row-tile kernel; 'tc': mma.sync row-tile kernel; 'tc5': tcgen05 or error.
        flag = config['xdr_fused_mlp'] if 'xdr_fused_mlp' in config else 'auto'
        self.fused_mlp_auto = flag == 'auto'
        self.fused_mlp_engine = 'tc5' if self.fused_mlp_auto else ops.fused_mlp_engine(flag)
        self.use_fused_mlp = self.fused_mlp_engine is not None
        self.source_latent_dim = config['source_embedding_size']
        self.target_latent_dim = config['target_embedding_size']
        self.reg_weight = config['reg_weight']
        self.map_func = config['mapping_function']
        # construction order == reference order (emcdr.py:57-71) so a seeded run draws the same initial weights
        if self.map_func == 'linear':
            self.mapping = nn.Linear(self.source_latent_dim, self.target_latent_dim, bias=False)
        else:
            assert config["mlp_hidden_size"] is not None
            dims = [self.source_latent_dim] + list(config["mlp_hidden_size"]) + [self.target_latent_dim]
            self.mapping = self.mlp_layers(dims)

        self.source_user_embedding = nn.Embedding(self.total_num_users, self.source_latent_dim)
        self.source_item_embedding = nn.Embedding(self.total_num_items, self.source_latent_dim)
        self.target_user_embedding = nn.Embedding(self.total_num_users, self.target_latent_dim)
        self.target_item_embedding = nn.Embedding(self.total_num_items, self.target_latent_dim)
        # The reference zero-fills the dead rows here (emcdr.py:73-78) and then overwrites them again with
        # xavier_normal_ (emcdr.py:84): no zero row survives, so only the init is reproduced.
        self.apply(xavier_normal_initialization)

    @staticmethod
    def mlp_layers(layer_dims):
        """Linear -> Tanh -> ... -> Linear, no activation after the last layer (emcdr.py:86-93)."""
        mods = []
        n = len(layer_dims) - 1
        for i in range(n):
            mods.append(nn.Linear(layer_dims[i], layer_dims[i + 1]))
            if i != n - 1:
                mods.append(nn.Tanh())
        return nn.Sequential(*mods)

    def set_phase(self, phase):
        self.phase = phase

    # ---- helpers ---------------------------------------------------------------------------------------------
    def _mapping_params(self):
        if isinstance(self.mapping, nn.Linear):
            return [self.mapping.weight], [None]
        lins = [m for m in self.mapping if isinstance(m, nn.Linear)]
        return [l.weight for l in lins], [l.bias for l in lins]

    def _apply_mapping(self, x):
        ws, bs = self._mapping_params()
        return ops.mlp_chain(x, ws, bs, _lib.ACT_TANH, _lib.ACT_NONE)

    def _tables(self, domain):
        if domain == 'source':
            return self.source_user_embedding.weight, self.source_item_embedding.weight
        return self.target_user_embedding.weight, self.target_item_embedding.weight

    # ---- scores ----------------------------------------------------------------------------------------------
    def source_forward(self, user, item):
        return ops.dot_score(*self._tables('source'), user, item)

    def target_forward(self, user, item):
        return ops.dot_score(*self._tables('target'), user, item)

    # ---- losses ----------------------------------------------------------------------------------------------
    def _domain_loss(self, interaction, domain):
        ut, it = self._tables(domain)
        uid = self.SOURCE_USER_ID if domain == 'source' else self.TARGET_USER_ID
        iid = self.SOURCE_ITEM_ID if domain == 'source' else self.TARGET_ITEM_ID
        if self.latent_factor_model == 'MF':
            label = interaction[self.SOURCE_LABEL if domain == 'source' else self.TARGET_LABEL]
            return ops.point_loss(ut, it, interaction[uid], interaction[iid], label, _lib.LOSS_MSE, self.reg_weight)
        neg = self.SOURCE_NEG_ITEM_ID if domain == 'source' else self.TARGET_NEG_ITEM_ID
        return ops.bpr_loss(ut, it, interaction[uid], interaction[iid], interaction[neg], self.reg_weight, self.bpr_gamma)

    def touched_rows(self, interaction):
        """(table parameter, ids) pairs whose rows ``calculate_loss(interaction)`` reads in the current phase -- what a
        row-sparse optimizer has to visit (trainer.RowSparseOptimizer)."""
        if self.phase == 'OVERLAP':
            idx = interaction[self.OVERLAP_ID].reshape(-1)
            if self.mode == 'overlap_users':
                return [(self.source_user_embedding.weight, idx), (self.target_user_embedding.weight, idx)]
            return [(self.source_item_embedding.weight, idx), (self.target_item_embedding.weight, idx)]
        domain = 'source' if self.phase == 'SOURCE' else 'target'
        ut, it = self._tables(domain)
        uid = self.SOURCE_USER_ID if domain == 'source' else self.TARGET_USER_ID
        iid = self.SOURCE_ITEM_ID if domain == 'source' else self.TARGET_ITEM_ID
        rows = [(ut, interaction[uid]), (it, interaction[iid])]
        if self.latent_factor_model != 'MF':
            rows.append((it, interaction[self.SOURCE_NEG_ITEM_ID if domain == 'source' else self.TARGET_NEG_ITEM_ID]))
        return rows

    def calculate_source_loss(self, interaction):
        return self._domain_loss(interaction, 'source')

    def calculate_target_loss(self, interaction):
        return self._domain_loss(interaction, 'target')

    def calculate_map_loss(self, interaction):
        idx = interaction[self.OVERLAP_ID]  # [b, 1] in the reference's dataloader (data/dataset.py:696)
        if self.mode == 'overlap_users':
            src, tgt = self.source_user_embedding.weight, self.target_user_embedding.weight
        else:
            src, tgt = self.source_item_embedding.weight, self.target_item_embedding.weight
        flat = idx.reshape(-1)
        ws, bs = self._mapping_params()
        dims = [ws[0].shape[1]] + [w.shape[0] for w in ws]
        if self.use_fused_mlp and ops.fused_mlp_supported(dims, self.fused_mlp_engine):
            # one kernel forward, one backward: gather -> MLP in shared memory -> MSE vs gathered target -> scatter
            return ops.fused_mlp_loss(0, 0, _lib.ACT_TANH, flat, None, None, (src, None, None, None, tgt), ws, bs,
                                      self.fused_mlp_engine)
        mapped = self._apply_mapping(ops.gather_rows(src, flat))
        return ops.mse_rows(mapped, tgt, flat)

    def calculate_loss(self, interaction):
        """Phase dispatch of emcdr.py:170-176: SOURCE -> source loss, OVERLAP -> map loss, anything else -> target."""
        if self.phase == 'SOURCE':
            return self.calculate_source_loss(interaction)
        elif self.phase == 'OVERLAP':
            return self.calculate_map_loss(interaction)
        else:
            return self.calculate_target_loss(interaction)

    def fused_step_spec(self):
        """What the trainer's persistent multi-step launch needs for the current phase (SOURCE / TARGET-like phases
        only; the OVERLAP phase has dense mapping parameters and runs batch by batch)."""
        if self.phase == 'OVERLAP':
            return None
        domain = 'source' if self.phase == 'SOURCE' else 'target'
        ut, it = self._tables(domain)
        uid = self.SOURCE_USER_ID if domain == 'source' else self.TARGET_USER_ID
        iid = self.SOURCE_ITEM_ID if domain == 'source' else self.TARGET_ITEM_ID
        if self.latent_factor_model == 'MF':
            return dict(user_tab=ut, item_tab=it, pairwise=False, loss_kind=_lib.LOSS_MSE, reg_weight=self.reg_weight,
                        fields=[uid, iid], label_field=self.SOURCE_LABEL if domain == 'source' else self.TARGET_LABEL)
        neg = self.SOURCE_NEG_ITEM_ID if domain == 'source' else self.TARGET_NEG_ITEM_ID
        return dict(user_tab=ut, item_tab=it, pairwise=True, reg_weight=self.reg_weight, gamma=self.bpr_gamma,
                    fields=[uid, iid, neg])

    # ---- inference -------------------------------------------------------------------------------------------
    def _mapped_user_e(self, user):
        mapped = self._apply_mapping(ops.gather_rows_raw(self.source_user_embedding.weight, user))
        return torch.where((user < self.overlapped_num_users).unsqueeze(1), mapped,
                           ops.gather_rows_raw(self.target_user_embedding.weight, user))

    def predict(self, interaction):
        if self.phase == 'SOURCE':
            return self.source_forward(interaction[self.SOURCE_USER_ID], interaction[self.SOURCE_ITEM_ID])
        if self.phase == 'TARGET':
            return self.target_forward(interaction[self.TARGET_USER_ID], interaction[self.TARGET_ITEM_ID])
        user = interaction[self.TARGET_USER_ID]
        item = interaction[self.TARGET_ITEM_ID]
        with torch.no_grad():
            if self.mode == 'overlap_users':
                mapped = self._apply_mapping(ops.gather_rows_raw(self.source_user_embedding.weight, user))
                return ops.select_dot(mapped, self.target_user_embedding.weight, user, self.overlapped_num_users,
                                      self.target_item_embedding.weight, item)
            mapped = self._apply_mapping(ops.gather_rows_raw(self.source_item_embedding.weight, item))
            return ops.select_dot(mapped, self.target_item_embedding.weight, item, self.overlapped_num_items,
                                  self.target_user_embedding.weight, user)

    def _full_sort_operands(self, interaction):
        """(user-side vectors [B, D], candidate item rows [n, D]) of full_sort_predict for the current phase
        (emcdr.py:208-233)."""
        with torch.no_grad():
            if self.phase == 'SOURCE':
                user_e = ops.gather_rows_raw(self.source_user_embedding.weight, interaction[self.SOURCE_USER_ID])
                w = self.source_item_embedding.weight
                all_item_e = torch.cat([w[:self.overlapped_num_items], w[self.target_num_items:]], dim=0)
            elif self.phase == 'TARGET':
                user_e = ops.gather_rows_raw(self.target_user_embedding.weight, interaction[self.TARGET_USER_ID])
                all_item_e = self.target_item_embedding.weight[:self.target_num_items]
            else:
                user = interaction[self.TARGET_USER_ID]
                if self.mode == 'overlap_users':
                    user_e = self._mapped_user_e(user)
                    all_item_e = self.target_item_embedding.weight[:self.target_num_items]
                else:
                    user_e = ops.gather_rows_raw(self.target_user_embedding.weight, user)
                    ov = self._apply_mapping(self.source_item_embedding.weight[:self.overlapped_num_items].contiguous())
                    all_item_e = torch.cat(
                        [ov, self.target_item_embedding.weight[self.overlapped_num_items:self.target_num_items]], dim=0)
            return user_e, all_item_e

    def full_sort_predict(self, interaction):
        """emcdr.py:208-233: the dense [B, n_items] score matrix, flattened (kept for drop-in compatibility; the GEMM is
        a plain library matmul).  Evaluation that only needs the best k items should call ``full_sort_topk``."""
        with torch.no_grad():
            user_e, all_item_e = self._full_sort_operands(interaction)
            return torch.matmul(user_e, all_item_e.transpose(0, 1)).view(-1)

    def full_sort_topk(self, interaction, k, hist_ptr=None, hist_ids=None, engine='mma'):
        """Fused form of ``full_sort_predict`` + recbole's full-sort masking (PAD column, per-user history) + ``topk``:
        one scoring kernel that never writes the [B, n_items] matrix.  Item positions are those of
        ``full_sort_predict``'s columns.  Returns (scores [B, k], positions [B, k])."""
        user_e, all_item_e = self._full_sort_operands(interaction)
        return ops.full_sort_topk(user_e, all_item_e.contiguous(), k, first_item=1, hist_ptr=hist_ptr, hist_ids=hist_ids,
                                  engine=engine)
