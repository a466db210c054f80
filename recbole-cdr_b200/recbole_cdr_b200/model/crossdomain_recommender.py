"""Base class of the model plugin API -- mirror of reference model/crossdomain_recommender.py:14-51 on top of the
members of ``recbole.model.abstract_recommender.AbstractRecommender`` [recbole-1.0.1] the trainer calls."""
import numpy as np
import torch.nn as nn

from ..utils import ModelType


class AbstractRecommender(nn.Module):
    """``calculate_loss`` / ``predict`` / ``full_sort_predict`` + ``other_parameter`` / ``load_other_parameter``."""

    def calculate_loss(self, interaction):
        raise NotImplementedError

    def predict(self, interaction):
        raise NotImplementedError

    def full_sort_predict(self, interaction):
        raise NotImplementedError

    def other_parameter(self):
        if hasattr(self, 'other_parameter_name'):
            return {key: getattr(self, key) for key in self.other_parameter_name}
        return dict()

    def load_other_parameter(self, para):
        if para is None:
            return
        for key, value in para.items():
            setattr(self, key, value)

    def __str__(self):
        params = sum(int(np.prod(p.size())) for p in self.parameters() if p.requires_grad)
        return super().__str__() + f'\nTrainable parameters: {params}'


class CrossDomainRecommender(AbstractRecommender):
    """Resolves field names and table sizes from the dataset exactly as the reference does
    (crossdomain_recommender.py:24-48).  The joint id layout those sizes describe is data/dataset.py:344-445:
    0 = [PAD]; [1, n_ov) overlapped; [n_ov, target_num) target-only; [target_num, total) source-only."""

    type = ModelType.CROSSDOMAIN

    def __init__(self, config, dataset):
        super().__init__()
        # per-domain field names and id-space sizes: SOURCE_USER_ID, SOURCE_ITEM_ID, SOURCE_NEG_ITEM_ID,
        # source_num_users, source_num_items and the same five for the target domain
        for domain, part in (('source', dataset.source_domain_dataset), ('target', dataset.target_domain_dataset)):
            tag = domain.upper()
            neg_prefix = config[f'{domain}_domain']['NEG_PREFIX']
            setattr(self, f'{tag}_USER_ID', part.uid_field)
            setattr(self, f'{tag}_ITEM_ID', part.iid_field)
            setattr(self, f'{tag}_NEG_ITEM_ID', neg_prefix + part.iid_field)
            setattr(self, f'{domain}_num_users', part.num(part.uid_field))
            setattr(self, f'{domain}_num_items', part.num(part.iid_field))
        # joint id space (every table is allocated with the TOTAL number of rows) and the overlapped prefix of it
        self.total_num_users, self.total_num_items = dataset.num_total_user, dataset.num_total_item
        self.overlapped_num_users, self.overlapped_num_items = dataset.num_overlap_user, dataset.num_overlap_item
        self.OVERLAP_ID = dataset.overlap_id_field
        self.device = config['device']

    def set_phase(self, phase):
        pass
