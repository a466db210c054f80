"""CPU oracle for the RecBole-CDR per-batch hot path.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file; the product package (``recbole-cdr_b200/``) never does.

What it is: a functional, plain-PyTorch fp32 (CPU) restatement of the arithmetic the reference
performs inside ``Model.calculate_loss`` / ``predict`` for EMCDR, CMF, CoNet, DTCDR(NeuMF) and BiTGCF,
plus the recbole-1.0.1 leaf losses those models call.  Gradients come from torch autograd on the
restated forward, i.e. they are the reference's gradients (dense ``[N, D]`` tensors; compare on the
touched rows).  Every function cites the reference ``file:line`` it follows (paths relative to
``/root/reference/recbole_cdr/``).

Parity pinning: the reference's own tests hold no golden values (``tests/test_model.py:10-11`` are
smoke runs).  This oracle is therefore pinned against outputs of the *unmodified reference model
classes executed in the build container*: ``oracle/make_golden.py`` imports them from
``/root/reference`` over ``oracle/recbole_shim`` and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those files.
The integer-side restatements (id layout, negative draw) live in ``oracle/sampler_oracle.py``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------------
# recbole-1.0.1 leaf losses (un-vendored dependency; semantics per SURVEY.md section 2)
# --------------------------------------------------------------------------------------------


def bpr_loss(pos_score: Tensor, neg_score: Tensor, gamma: float = 1e-10) -> Tensor:
    """recbole.model.loss.BPRLoss: -(log(gamma + sigmoid(pos - neg))).mean()  (call sites emcdr.py:54,128,151)."""
    return -(torch.log(gamma + torch.sigmoid(pos_score - neg_score))).mean()


def emb_loss(*embeddings: Tensor) -> Tensor:
    """recbole.model.loss.EmbLoss(norm=2), require_pow=False: (sum_k ||E_k||_F) / E_last.shape[0], shape [1].

    Call sites: emcdr.py:119,129,142,152; cmf.py:94-98; bitgcf.py:233,245.
    """
    acc = torch.zeros(1, dtype=embeddings[-1].dtype, device=embeddings[-1].device)
    for e in embeddings:
        acc = acc + torch.norm(e, p=2)
    return acc / embeddings[-1].shape[0]


def bce_loss(prob: Tensor, label: Tensor) -> Tensor:
    """torch.nn.BCELoss (mean; log clamped at -100) as used by cmf.py:45, conet.py:63, dtcdr.py:103, bitgcf.py:67."""
    return F.binary_cross_entropy(prob, label)


def mse_loss(pred: Tensor, target: Tensor) -> Tensor:
    """torch.nn.MSELoss (mean) as used by emcdr.py:50,81."""
    return F.mse_loss(pred, target)


# --------------------------------------------------------------------------------------------
# A1/A2: gather + dot-product score
# --------------------------------------------------------------------------------------------


def gather_rows(table: Tensor, idx: Tensor) -> Tensor:
    """nn.Embedding.__call__ == table[idx]  (emcdr.py:99-100; bit-exact contract)."""
    return table[idx]


def dot_score(user_tab: Tensor, item_tab: Tensor, user: Tensor, item: Tensor) -> Tensor:
    """EMCDR.source_forward / target_forward, emcdr.py:98-108: (Eu[u] * Ei[i]).sum(dim=1)."""
    return (user_tab[user] * item_tab[item]).sum(dim=1)


# --------------------------------------------------------------------------------------------
# A3: EMCDR per-domain recommendation loss
# --------------------------------------------------------------------------------------------


def emcdr_bpr_loss(user_tab: Tensor, item_tab: Tensor, user: Tensor, pos_item: Tensor, neg_item: Tensor,
                   reg_weight: float) -> Tensor:
    """EMCDR.calculate_source_loss / calculate_target_loss, BPR branch: emcdr.py:121-130 / :144-153.

    BPRLoss(s(u,i+), s(u,i-)) + reg_weight * EmbLoss(Eu[u], Ei[i+]); result has shape [1].
    """
    pos = dot_score(user_tab, item_tab, user, pos_item)
    neg = dot_score(user_tab, item_tab, user, neg_item)
    return bpr_loss(pos, neg) + reg_weight * emb_loss(user_tab[user], item_tab[pos_item])


def emcdr_mf_loss(user_tab: Tensor, item_tab: Tensor, user: Tensor, item: Tensor, label: Tensor,
                  reg_weight: float) -> Tensor:
    """EMCDR MF branch: emcdr.py:111-120 / :134-143.  MSELoss(s(u,i), label) + reg_weight * EmbLoss(...)."""
    pred = dot_score(user_tab, item_tab, user, item)
    return mse_loss(pred, label) + reg_weight * emb_loss(user_tab[user], item_tab[item])


# --------------------------------------------------------------------------------------------
# A4/A6: EMCDR mapping function, map loss and fused predict
# --------------------------------------------------------------------------------------------


def emcdr_mapping(x: Tensor, weights: Sequence[Tensor], biases: Sequence[Optional[Tensor]]) -> Tensor:
    """EMCDR.mapping: Linear(bias=False) (emcdr.py:58-59) or Linear->Tanh->...->Linear with NO final Tanh
    (emcdr.py:86-93)."""
    n = len(weights)
    for k in range(n):
        x = F.linear(x, weights[k], biases[k])
        if k != n - 1:
            x = torch.tanh(x)
    return x


def emcdr_map_loss(src_tab: Tensor, tgt_tab: Tensor, idx: Tensor, weights: Sequence[Tensor],
                   biases: Sequence[Optional[Tensor]]) -> Tensor:
    """EMCDR.calculate_map_loss, emcdr.py:156-168: MSELoss(mapping(Es[idx]), Et[idx]); target NOT detached.

    ``idx`` keeps the reference's [b, 1] shape (dataset.py:696) -- the gathers are then [b, 1, D] and the
    mean is over the same b*D elements.
    """
    return mse_loss(emcdr_mapping(src_tab[idx], weights, biases), tgt_tab[idx])


def emcdr_predict_overlap_users(src_user_tab: Tensor, tgt_user_tab: Tensor, tgt_item_tab: Tensor, user: Tensor,
                                item: Tensor, n_overlap_users: int, weights, biases) -> Tensor:
    """EMCDR.predict, phase not SOURCE/TARGET, mode overlap_users: emcdr.py:192-199,205."""
    mapped = emcdr_mapping(src_user_tab[user], weights, biases)
    sel = (user < n_overlap_users).unsqueeze(1)
    user_e = torch.where(sel, mapped, tgt_user_tab[user])
    return (user_e * tgt_item_tab[item]).sum(dim=1)


def emcdr_predict_overlap_items(tgt_user_tab: Tensor, src_item_tab: Tensor, tgt_item_tab: Tensor, user: Tensor,
                                item: Tensor, n_overlap_items: int, weights, biases) -> Tensor:
    """EMCDR.predict, mode overlap_items: emcdr.py:200-205."""
    mapped = emcdr_mapping(src_item_tab[item], weights, biases)
    sel = (item < n_overlap_items).unsqueeze(1)
    item_e = torch.where(sel, mapped, tgt_item_tab[item])
    return (tgt_user_tab[user] * item_e).sum(dim=1)


# --------------------------------------------------------------------------------------------
# A16: CMF
# --------------------------------------------------------------------------------------------


def cmf_loss(user_tab: Tensor, item_tab: Tensor, su: Tensor, si: Tensor, sl: Tensor, tu: Tensor, ti: Tensor,
             tl: Tensor, alpha: float, lamda: float, gamma: float) -> Tensor:
    """CMF.calculate_loss, cmf.py:81-99: alpha*(BCE_s + lambda*Emb_s) + (1-alpha)*(BCE_t + gamma*Emb_t);
    one shared user table and one shared item table; p = sigmoid(dot) (cmf.py:75-79)."""
    p_s = torch.sigmoid(dot_score(user_tab, item_tab, su, si))
    p_t = torch.sigmoid(dot_score(user_tab, item_tab, tu, ti))
    loss_s = bce_loss(p_s, sl) + lamda * emb_loss(user_tab[su], item_tab[si])
    loss_t = bce_loss(p_t, tl) + gamma * emb_loss(user_tab[tu], item_tab[ti])
    return loss_s * alpha + loss_t * (1 - alpha)


# --------------------------------------------------------------------------------------------
# A7/A8: CoNet
# --------------------------------------------------------------------------------------------


def conet_towers(tabs: Dict[str, Tensor], user: Tensor, item: Tensor, p: Dict[str, List[Tensor]],
                 overlap_users: bool, n_overlap: int) -> Tuple[Tensor, Tensor]:
    """The shared body of CoNet.source_forward / target_forward, conet.py:105-142 / :144-181.

    x_s = [Es_u[u] | Es_i[i]], x_t = [Et_u[u] | Et_i[i]];  per layer l:
      h_s = W_s x_s + b_s + m * (x_t H_l),  h_t = W_t x_t + b_t + m * (x_s H_l),  H_l = crossparas[l].weight.T,
      m = (u < n_ov_users) or (i < n_ov_items);  ReLU.  Returns (sigmoid(out_s), sigmoid(out_t)), each [B].
    ``p``: lists ``ws, bs, wt, bt, h`` per layer and ``out_s_w, out_s_b, out_t_w, out_t_b``.
    """
    x_s = torch.cat([tabs['source_user'][user], tabs['source_item'][item]], dim=1)
    x_t = torch.cat([tabs['target_user'][user], tabs['target_item'][item]], dim=1)
    m = ((user < n_overlap) if overlap_users else (item < n_overlap)).to(x_s.dtype).unsqueeze(1)
    for l in range(len(p['ws'])):
        cross = p['h'][l].t()
        h_s = torch.relu(F.linear(x_s, p['ws'][l], p['bs'][l]) + m * (x_t @ cross))
        h_t = torch.relu(F.linear(x_t, p['wt'][l], p['bt'][l]) + m * (x_s @ cross))
        x_s, x_t = h_s, h_t
    out_s = torch.sigmoid(F.linear(x_s, p['out_s_w'], p['out_s_b'])).squeeze(-1)
    out_t = torch.sigmoid(F.linear(x_t, p['out_t_w'], p['out_t_b'])).squeeze(-1)
    return out_s, out_t


def conet_loss(tabs, p, su, si, sl, tu, ti, tl, overlap_users: bool, n_overlap: int) -> Tensor:
    """CoNet.calculate_loss, conet.py:183-203: BCE(source tower on source batch) + BCE(target tower on target
    batch) + sum_l ||H_l||_F  (reg_weight is read at conet.py:53 but never applied)."""
    p_s, _ = conet_towers(tabs, su, si, p, overlap_users, n_overlap)
    _, p_t = conet_towers(tabs, tu, ti, p, overlap_users, n_overlap)
    reg = sum(torch.norm(h) for h in p['h'])
    return bce_loss(p_s, sl) + bce_loss(p_t, tl) + reg


def conet_predict(tabs, p, user, item) -> Tensor:
    """CoNet.predict, conet.py:205-220: target tower WITHOUT cross terms; returns [B, 1]."""
    x = torch.cat([tabs['target_user'][user], tabs['target_item'][item]], dim=1)
    for l in range(len(p['wt'])):
        x = torch.relu(F.linear(x, p['wt'][l], p['bt'][l]))
    return torch.sigmoid(F.linear(x, p['out_t_w'], p['out_t_b']))


# --------------------------------------------------------------------------------------------
# A14/A15: DTCDR (NeuMF base model)
# --------------------------------------------------------------------------------------------


def dtcdr_neumf_forward(tabs, user: Tensor, item: Tensor, mlp_w: Sequence[Tensor], mlp_b: Sequence[Tensor],
                        out_w: Tensor, out_b: Tensor) -> Tensor:
    """DTCDR.neumf_forward, dtcdr.py:112-125 (dropout_prob = 0): u = max(Es_u[u], Et_u[u]), i likewise;
    sigmoid(Linear(MLPLayers([u|i]))) with ReLU after EVERY MLP layer (recbole MLPLayers)."""
    u = torch.maximum(tabs['source_user'][user], tabs['target_user'][user])
    i = torch.maximum(tabs['source_item'][item], tabs['target_item'][item])
    x = torch.cat((u, i), -1)
    for w, b in zip(mlp_w, mlp_b):
        x = torch.relu(F.linear(x, w, b))
    return torch.sigmoid(F.linear(x, out_w, out_b)).squeeze(-1)


def dtcdr_loss(tabs, p, su, si, sl, tu, ti, tl, alpha: float) -> Tensor:
    """DTCDR.calculate_loss, NeuMF branch, dtcdr.py:177-191: alpha*BCE_s + (1-alpha)*BCE_t."""
    o_s = dtcdr_neumf_forward(tabs, su, si, p['s_mlp_w'], p['s_mlp_b'], p['s_out_w'], p['s_out_b'])
    o_t = dtcdr_neumf_forward(tabs, tu, ti, p['t_mlp_w'], p['t_mlp_b'], p['t_out_w'], p['t_out_b'])
    return bce_loss(o_s, sl) * alpha + bce_loss(o_t, tl) * (1 - alpha)


# --------------------------------------------------------------------------------------------
# A9-A13: BiTGCF
# --------------------------------------------------------------------------------------------


def bitgcf_norm_adj(rows: np.ndarray, cols: np.ndarray, n_users: int, n_items: int) -> Tensor:
    """BiTGCF.get_norm_adj_mat, bitgcf.py:92-116, built through COO instead of the private dok_matrix._update
    (absent in SciPy >= 1.13): A = [[0, R], [R^T, 0]] over n_users+n_items nodes with unit entries (duplicates
    collapse to 1, as the reference's dict does), D_ii = rowcount(A > 0) + 1e-7, L = D^-1/2 A D^-1/2 (fp32 COO).
    """
    n = n_users + n_items
    edges = np.unique(np.stack([rows.astype(np.int64), cols.astype(np.int64)], 1), axis=0)
    r = np.concatenate([edges[:, 0], edges[:, 1] + n_users])
    c = np.concatenate([edges[:, 1] + n_users, edges[:, 0]])
    deg = np.bincount(r, minlength=n).astype(np.float64) + 1e-7
    dinv = np.power(deg, -0.5)
    # scipy: D * A * D with A float32 and D float64 -> float64 product, then torch.FloatTensor(L.data)
    val = (dinv[r] * 1.0 * dinv[c]).astype(np.float32)
    order = np.lexsort((c, r))
    idx = torch.from_numpy(np.stack([r[order], c[order]]))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(val[order]), (n, n)).coalesce()


def bitgcf_graph_layer(adj: Tensor, e: Tensor) -> Tensor:
    """BiTGCF.graph_layer, bitgcf.py:130-135 with drop_rate = 0: S = L.E; E' = E + S + E*S."""
    s = torch.sparse.mm(adj, e)
    return e + (s + e * s)


def bitgcf_transfer_layer(es: Tensor, et: Tensor, n_users: int, n_items: int, n_ov_users: int, n_ov_items: int,
                          lam_s: float, lam_t: float, deg: Dict[str, Tensor]) -> Tuple[Tensor, Tensor]:
    """BiTGCF.transfer_layer, bitgcf.py:137-172.  Rows < n_ov (users and items separately) become
    0.5*[(lam*E_own + (1-lam)*E_other) + (d_s*E_s + d_t*E_t)/(d_s + d_t + 1e-7)]; all other rows unchanged.
    ``deg``: 'su','tu' [n_users,1], 'si','ti' [n_items,1] fp32 degree counts (bitgcf.py:79-82)."""
    su, si = torch.split(es, [n_users, n_items])
    tu, ti = torch.split(et, [n_users, n_items])

    def mix(a_s, a_t, d_s, d_t, n_ov):
        lam_src = lam_s * a_s + (1 - lam_s) * a_t
        lam_tgt = lam_t * a_t + (1 - lam_t) * a_s
        lap = (d_s * a_s + d_t * a_t) / (d_s + d_t + 1e-7)
        new_s = torch.cat([(lam_src[:n_ov] + lap[:n_ov]) / 2, a_s[n_ov:]], dim=0)
        new_t = torch.cat([(lam_tgt[:n_ov] + lap[:n_ov]) / 2, a_t[n_ov:]], dim=0)
        return new_s, new_t

    nsu, ntu = mix(su, tu, deg['su'], deg['tu'], n_ov_users)
    nsi, nti = mix(si, ti, deg['si'], deg['ti'], n_ov_items)
    return torch.cat([nsu, nsi], 0), torch.cat([ntu, nti], 0)


def bitgcf_forward(tabs, adj_s: Tensor, adj_t: Tensor, n_layers: int, connect_way: str, n_users: int, n_items: int,
                   n_ov_users: int, n_ov_items: int, lam_s: float, lam_t: float, deg) -> Tuple[Tensor, ...]:
    """BiTGCF.forward, bitgcf.py:174-205: per layer graph_layer (both domains) -> transfer_layer ->
    F.normalize(p=2, dim=1) -> append; combine by 'concat' or 'mean'; split users/items."""
    es = torch.cat([tabs['source_user'], tabs['source_item']], 0)
    et = torch.cat([tabs['target_user'], tabs['target_item']], 0)
    ls, lt = [es], [et]
    for _ in range(n_layers):
        es = bitgcf_graph_layer(adj_s, es)
        et = bitgcf_graph_layer(adj_t, et)
        es, et = bitgcf_transfer_layer(es, et, n_users, n_items, n_ov_users, n_ov_items, lam_s, lam_t, deg)
        ls.append(F.normalize(es, p=2, dim=1))
        lt.append(F.normalize(et, p=2, dim=1))
    if connect_way == 'concat':
        fs, ft = torch.cat(ls, 1), torch.cat(lt, 1)
    else:
        fs, ft = torch.stack(ls, 1).mean(1), torch.stack(lt, 1).mean(1)
    su, si = torch.split(fs, [n_users, n_items])
    tu, ti = torch.split(ft, [n_users, n_items])
    return su, si, tu, ti


def bitgcf_loss(tabs, adj_s, adj_t, su, si, sl, tu, ti, tl, *, n_layers, connect_way, n_users, n_items, n_ov_users,
                n_ov_items, lam_s, lam_t, deg, reg_weight) -> Tuple[Tensor, Tensor]:
    """BiTGCF.calculate_loss, bitgcf.py:207-250: per domain BCE(sigmoid(dot of propagated rows)) +
    reg_weight*EmbLoss(ego rows); returns the TUPLE (source_loss, target_loss), each shape [1]."""
    fsu, fsi, ftu, fti = bitgcf_forward(tabs, adj_s, adj_t, n_layers, connect_way, n_users, n_items, n_ov_users,
                                        n_ov_items, lam_s, lam_t, deg)
    p_s = torch.sigmoid((fsu[su] * fsi[si]).sum(1))
    loss_s = bce_loss(p_s, sl) + reg_weight * emb_loss(tabs['source_user'][su], tabs['source_item'][si])
    p_t = torch.sigmoid((ftu[tu] * fti[ti]).sum(1))
    loss_t = bce_loss(p_t, tl) + reg_weight * emb_loss(tabs['target_user'][tu], tabs['target_item'][ti])
    return loss_s, loss_t


# --------------------------------------------------------------------------------------------
# helpers shared by tests and the CPU-baseline leg of bench.py
# --------------------------------------------------------------------------------------------


def xavier_normal_table(n_rows: int, dim: int, generator: Optional[torch.Generator] = None,
                        std: Optional[float] = None) -> Tensor:
    """recbole xavier_normal_initialization on an [N, D] table: N(0, 2/(N+D))  (emcdr.py:84)."""
    s = math.sqrt(2.0 / (n_rows + dim)) if std is None else std
    return torch.randn(n_rows, dim, generator=generator) * s


def grads_of(loss: Tensor, params: Sequence[Tensor]) -> List[Tensor]:
    """Dense autograd gradients (what loss.backward() leaves in .grad in the reference's trainer step)."""
    out = torch.autograd.grad(loss.sum(), list(params), allow_unused=True)
    return [torch.zeros_like(p) if g is None else g for p, g in zip(params, out)]


class DenseTrainerStep:
    """Restatement of one recbole ``Trainer._train_epoch`` iteration [recbole-1.0.1] as driven by
    trainer/trainer.py:59-73: zero_grad -> calculate_loss -> backward (dense [N, D] grads via
    embedding_dense_backward) -> dense Adam.  Used as the timed CPU baseline (cost structure of the reference)."""

    def __init__(self, params: Sequence[Tensor], lr: float = 1e-3, with_optimizer: bool = True):
        self.params = [torch.nn.Parameter(p) for p in params]
        self.opt = torch.optim.Adam(self.params, lr=lr, weight_decay=0.0) if with_optimizer else None

    def step(self, loss_fn, backward: bool = True) -> float:
        if self.opt is not None:
            self.opt.zero_grad()
        else:
            for p in self.params:
                p.grad = None
        loss = loss_fn(*self.params)
        loss = sum(loss) if isinstance(loss, tuple) else loss
        if backward:
            loss.sum().backward()
            if self.opt is not None:
                self.opt.step()
        return float(loss.sum().item())
