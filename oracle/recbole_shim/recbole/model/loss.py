"""[recbole-1.0.1] leaf losses used on the RecBole-CDR hot path."""
import torch
import torch.nn as nn


class BPRLoss(nn.Module):
    """-(log(gamma + sigmoid(pos - neg))).mean(); gamma = 1e-10; NOT logsigmoid."""

    def __init__(self, gamma=1e-10):
        super().__init__()
        self.gamma = gamma

    def forward(self, pos_score, neg_score):
        return -torch.log(self.gamma + torch.sigmoid(pos_score - neg_score)).mean()


class EmbLoss(nn.Module):
    """require_pow=False (default): (sum_k ||E_k||_F) / E_last.shape[0]; returns shape [1]."""

    def __init__(self, norm=2):
        super().__init__()
        self.norm = norm

    def forward(self, *embeddings, require_pow=False):
        emb_loss = torch.zeros(1).to(embeddings[-1].device)
        if require_pow:
            for embedding in embeddings:
                emb_loss += torch.pow(input=torch.norm(embedding, p=self.norm), exponent=self.norm)
            emb_loss /= embeddings[-1].shape[0]
            emb_loss /= self.norm
            return emb_loss
        for embedding in embeddings:
            emb_loss += torch.norm(embedding, p=self.norm)
        emb_loss /= embeddings[-1].shape[0]
        return emb_loss


class RegLoss(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, parameters):
        reg_loss = None
        for W in parameters:
            reg_loss = W.norm(2) if reg_loss is None else reg_loss + W.norm(2)
        return reg_loss
