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
        src, tgt = dataset.source_domain_dataset, dataset.target_domain_dataset
        self.SOURCE_USER_ID = src.uid_field
        self.SOURCE_ITEM_ID = src.iid_field
        self.SOURCE_NEG_ITEM_ID = config['source_domain']['NEG_PREFIX'] + self.SOURCE_ITEM_ID
        self.source_num_users = src.num(self.SOURCE_USER_ID)
        self.source_num_items = src.num(self.SOURCE_ITEM_ID)

        self.TARGET_USER_ID = tgt.uid_field
        self.TARGET_ITEM_ID = tgt.iid_field
        self.TARGET_NEG_ITEM_ID = config['target_domain']['NEG_PREFIX'] + self.TARGET_ITEM_ID
        self.target_num_users = tgt.num(self.TARGET_USER_ID)
        self.target_num_items = tgt.num(self.TARGET_ITEM_ID)

        self.total_num_users = dataset.num_total_user
        self.total_num_items = dataset.num_total_item
        self.overlapped_num_users = dataset.num_overlap_user
        self.overlapped_num_items = dataset.num_overlap_item
        self.OVERLAP_ID = dataset.overlap_id_field

        self.device = config['device']

    def set_phase(self, phase):
        pass
