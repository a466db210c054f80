"""CPU oracle for the five remaining RecBole-CDR models.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Functional plain-PyTorch fp32 restatements of ``calculate_loss`` of CLFM, DeepAPF, SSCDR, NATR and DCDCSR; parameters come
in as a ``{state_dict key: tensor}`` dict ``P`` with the reference's own names, gradients from autograd.  Every function cites
the reference ``file:line`` it follows (paths relative to ``/root/reference/recbole_cdr/model/cross_domain_recommender/``).
Pinned against outputs of the UNMODIFIED reference classes: ``oracle/make_golden_f4.py`` -> ``tests/golden/f4_*.npz``,
checked by ``tests/test_oracle_f4.py``.  Only ``tests/`` may import this module."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .cdr_oracle import bce_loss, bpr_loss, emb_loss

Tensor = torch.Tensor


# ---------------------------------------------------------------------------------------------------------------- CLFM
def clfm_forward(P: Dict[str, Tensor], domain: str, user: Tensor, item: Tensor) -> Tensor:
    """CLFM.source_forward / target_forward, clfm.py:70-98: sigmoid(<[shared(u) | only(u)], item row>)."""
    u = P[f'{domain}_user_embedding.weight'][user]
    parts = []
    if 'shared_linear.weight' in P:
        parts.append(F.linear(u, P['shared_linear.weight']))
    if f'{domain}_only_linear.weight' in P:
        parts.append(F.linear(u, P[f'{domain}_only_linear.weight']))
    return torch.sigmoid((torch.cat(parts, dim=1) * P[f'{domain}_item_embedding.weight'][item]).sum(dim=1))


def clfm_loss(P, su, si, sl, tu, ti, tl, alpha: float, reg_weight: float) -> Tensor:
    """CLFM.calculate_loss, clfm.py:100-122."""
    def term(domain, u, i, l):
        reg = emb_loss(P[f'{domain}_user_embedding.weight'][u], P[f'{domain}_item_embedding.weight'][i])
        return bce_loss(clfm_forward(P, domain, u, i), l) + reg_weight * reg
    return term('source', su, si, sl) * alpha + term('target', tu, ti, tl) * (1 - alpha)


# ------------------------------------------------------------------------------------------------------------- DeepAPF
def deepapf_forward(P, domain: str, user: Tensor, item: Tensor, overlap_users: bool, n_overlap: int) -> Tensor:
    """DeepAPF.source_forward / target_forward, deepapf.py:68-146.  The shared and the domain-only row of the overlapped
    side get attention logits from a two-layer MLP on their product with the other side's row; the shared logit is masked
    with -1e31 where id > n_overlap (strict, deepapf.py:73); softmax over the two; predict_layer on (mixed * other).
    The item MLP is registered as ``seq`` in the reference's named_parameters() (deepapf.py:60-62)."""
    if overlap_users:
        share, only = P['share_user_embedding.weight'][user], P[f'{domain}_user_embedding.weight'][user]
        other, mask, pre = P[f'{domain}_item_embedding.weight'][item], (user > n_overlap).unsqueeze(-1), 'user_mlp'
    else:
        share, only = P['share_item_embedding.weight'][item], P[f'{domain}_item_embedding.weight'][item]
        other, mask, pre = P[f'{domain}_user_embedding.weight'][user], (item > n_overlap).unsqueeze(-1), 'seq'

    def mlp(x):
        return F.linear(torch.relu(F.linear(x, P[f'{pre}.0.weight'], P[f'{pre}.0.bias'])), P[f'{pre}.2.weight'])

    a_share = mlp(share * other).masked_fill(mask, -1e31)
    alpha = torch.softmax(torch.cat([a_share, mlp(only * other)], dim=1), dim=1)
    mixed = alpha[:, 0:1] * share + alpha[:, 1:2] * only
    return torch.sigmoid(F.linear(mixed * other, P['predict_layer.weight'])).squeeze(-1)


def deepapf_loss(P, su, si, sl, tu, ti, tl, overlap_users: bool, n_overlap: int) -> Tensor:
    """DeepAPF.calculate_loss, deepapf.py:158-175: BCE(source) + BCE(target)."""
    return bce_loss(deepapf_forward(P, 'source', su, si, overlap_users, n_overlap), sl) + \
        bce_loss(deepapf_forward(P, 'target', tu, ti, overlap_users, n_overlap), tl)


# --------------------------------------------------------------------------------------------------------------- SSCDR
def sscdr_normalize(e: Tensor) -> Tensor:
    """SSCDR.embedding_normalize, sscdr.py:124-129: divide by the SQUARED length where that exceeds 1."""
    sq = torch.sum(e ** 2, dim=1, keepdim=True)
    return e / torch.where(sq > 1, sq, torch.ones_like(sq))


def sscdr_mapping(P, x: Tensor) -> Tensor:
    """recbole MLPLayers(activation='tanh'): Tanh after EVERY Linear (sscdr.py:48-49); keys mapping_layer.mlp_layers.{1,4,..}."""
    k = 1
    while f'mapping_layer.mlp_layers.{k}.weight' in P:
        x = torch.tanh(F.linear(x, P[f'mapping_layer.mlp_layers.{k}.weight'], P[f'mapping_layer.mlp_layers.{k}.bias']))
        k += 3
    return x


def sscdr_triplet(anchor, pos, neg, margin: float) -> Tensor:
    """nn.TripletMarginLoss(margin) on length-clipped rows (sscdr.py:70, 143-145)."""
    return F.triplet_margin_loss(sscdr_normalize(anchor), sscdr_normalize(pos), sscdr_normalize(neg), margin=margin)


def sscdr_rec_loss(P, domain: str, user, pos, neg, margin: float) -> Tensor:
    """SSCDR.calculate_source_loss / calculate_target_loss, sscdr.py:134-160."""
    ut, it = P[f'{domain}_user_embedding.weight'], P[f'{domain}_item_embedding.weight']
    return sscdr_triplet(ut[user], it[pos], it[neg], margin)


def sscdr_map_loss(P, idx, pos, neg, overlap_users: bool, margin: float, lamda: float) -> Tensor:
    """SSCDR.calculate_map_loss, sscdr.py:162-187, with the host-side draws (pos, neg) of SSCDR.sample given."""
    side, other = ('user', 'item') if overlap_users else ('item', 'user')
    src, tgt = P[f'source_{side}_embedding.weight'][idx], P[f'target_{side}_embedding.weight'][idx]
    oth = P[f'source_{other}_embedding.weight']
    loss_s = F.mse_loss(sscdr_mapping(P, src), tgt)
    loss_u = sscdr_triplet(tgt, sscdr_mapping(P, oth[pos]), sscdr_mapping(P, oth[neg]), margin)
    return loss_s + lamda * loss_u


# ---------------------------------------------------------------------------------------------------------------- NATR
def natr_phase1_loss(P, user, item, label) -> Tensor:
    """NATR.calculate_phase1_loss, natr.py:103-115."""
    s = (P['source_user_embedding.weight'][user] * P['source_item_embedding.weight'][item]).sum(dim=1)
    return bce_loss(torch.sigmoid(s), label)


def natr_phase2_forward(P, user, item, overlap_items: bool, history: Tensor, mask_mat: Tensor) -> Tensor:
    """NATR.phase2_forward, natr.py:117-160.  ``history`` / ``mask_mat``: the target-domain history matrix truncated to
    max_inter_length and its validity mask (natr.py:85-101), rows indexed by user (overlap_items) or item (overlap_users)."""
    user_e, item_e = P['target_user_embedding.weight'][user], P['target_item_embedding.weight'][item]
    if overlap_items:
        key, src, pu, qi = user, P['source_item_embedding.weight'], user_e, item_e
    else:
        key, src, pu, qi = item, P['source_user_embedding.weight'], item_e, user_e
    bias = torch.where(mask_mat[key].bool(), 0., -10000.0)
    h = F.linear(src[history[key]], P['transfer_layer.weight'], P['transfer_layer.bias'])            # [B, H, D]
    att = F.linear(torch.relu(pu.unsqueeze(1) * h), P['unit_attention_layer.weight'], P['unit_attention_layer.bias']).squeeze(2)
    su = torch.bmm(torch.softmax(att + bias, dim=1).unsqueeze(1), h).squeeze(1)
    dw, db = P['domain_attention_layer.weight'], P['domain_attention_layer.bias']
    b_s, b_p = F.linear(torch.relu(su * qi), dw, db), F.linear(torch.relu(pu * qi), dw, db)
    beta_s = torch.exp(b_s) / (torch.exp(b_s) + torch.exp(b_p))
    return torch.sigmoid(((beta_s * su + (1 - beta_s) * pu) * qi).sum(dim=1))


def natr_phase2_loss(P, user, item, label, overlap_items: bool, history, mask_mat, reg_weight: float) -> Tensor:
    """NATR.calculate_phase2_loss, natr.py:162-172: BCE + reg_weight * RegLoss (sum of the parameters' 2-norms)."""
    reg = sum(P[k].norm(2) for k in ('target_user_embedding.weight', 'target_item_embedding.weight', 'transfer_layer.weight',
                                     'unit_attention_layer.weight', 'domain_attention_layer.weight'))
    return bce_loss(natr_phase2_forward(P, user, item, overlap_items, history, mask_mat), label) + reg_weight * reg


# -------------------------------------------------------------------------------------------------------------- DCDCSR
def dcdcsr_rec_loss(user_tab: Tensor, item_tab: Tensor, user, pos, neg) -> Tensor:
    """DCDCSR.calculate_rec_loss (BPR), dcdcsr.py:120-134."""
    u = user_tab[user]
    return bpr_loss((u * item_tab[pos]).sum(dim=1), (u * item_tab[neg]).sum(dim=1))


def dcdcsr_benchmark(src_ov: Tensor, tgt: Tensor, pop_s: Tensor, pop_t: Tensor, k: int) -> Tensor:
    """DCDCSR.build_unit_benchmark_embedding, dcdcsr.py:136-159, unit by unit exactly as the reference loops.
    ``src_ov``: the overlapped source rows [n_ov, D]; ``tgt``: all target-side rows [n_total, D]."""
    n_ov, n_total = src_ov.shape[0], tgt.shape[0]
    out = torch.empty_like(tgt)
    for idx in range(n_ov):
        den = pop_s[idx] + pop_t[idx]
        den = den if den != 0 else 1
        a_s = pop_s[idx] / den
        out[idx] = a_s * tgt[idx] + (1 - a_s) * src_ov[idx]
    for idx in range(n_ov, n_total):
        sim, index = torch.topk(src_ov @ tgt[idx], k=k, dim=0)
        sn = torch.mean(pop_s[index])
        beta = sn / (sn + pop_t[idx])
        tot = torch.sum(sim) if torch.sum(sim) > 0 else 1
        out[idx] = (1 - beta) * tgt[idx] + beta * ((sim.unsqueeze(0) @ src_ov[index]).squeeze(0) / tot)
    return out.detach()


def dcdcsr_maxmin(w: Tensor):
    """DCDCSR.maxmin_normalize, dcdcsr.py:172-177."""
    mn, mx = torch.amin(w, dim=1, keepdim=True), torch.amax(w, dim=1, keepdim=True)
    mean = (mx + mn) / 2
    return (w - mean) / (mx - mean), mean, mx


def dcdcsr_mapping(P, x: Tensor) -> Tensor:
    k = 1
    while f'mapping_mlp_layers.mlp_layers.{k}.weight' in P:
        x = torch.tanh(F.linear(x, P[f'mapping_mlp_layers.mlp_layers.{k}.weight'], P[f'mapping_mlp_layers.mlp_layers.{k}.bias']))
        k += 3
    return x


def dcdcsr_map_loss(P, tgt_tab: Tensor, benchmark: Tensor, sampled: Tensor) -> Tensor:
    """DCDCSR.calculate_unit_map_loss, dcdcsr.py:179-188, with the sampled unit ids given."""
    rows, _, _ = dcdcsr_maxmin(tgt_tab[sampled])
    bench, _, _ = dcdcsr_maxmin(benchmark[sampled])
    return F.mse_loss(dcdcsr_mapping(P, rows), bench)
