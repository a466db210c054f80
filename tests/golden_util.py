"""Loading helpers for tests/golden/*.npz (written by oracle/make_golden.py from the reference's own classes)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))

    def t(self, key):
        return torch.from_numpy(self.z[key])

    def param(self, name):
        return self.t('param/' + name)

    def grad(self, name):
        return self.t('grad/' + name)

    def batch(self, name):
        return self.t('batch/' + name)

    def meta(self, name):
        return self.z['meta/' + name].item()

    def has(self, key):
        return key in self.z.files

    def param_names(self):
        return [k[len('param/'):] for k in self.z.files if k.startswith('param/')]

    def losses(self):
        return [self.t(k) for k in sorted(k for k in self.z.files if k.startswith('loss'))]

    def tables(self):
        return {k: self.param(f'{k}_embedding.weight') for k in
                ('source_user', 'source_item', 'target_user', 'target_item')}


def emcdr_mapping_params(g):
    ws, bs = [], []
    if g.has('param/mapping.weight'):
        return [g.param('mapping.weight')], [None], ['mapping.weight'], [None]
    wn, bn = [], []
    k = 0
    while g.has(f'param/mapping.{k}.weight'):
        ws.append(g.param(f'mapping.{k}.weight'))
        bs.append(g.param(f'mapping.{k}.bias'))
        wn.append(f'mapping.{k}.weight')
        bn.append(f'mapping.{k}.bias')
        k += 2
    return ws, bs, wn, bn


def conet_params(g):
    p = {'ws': [], 'bs': [], 'wt': [], 'bt': [], 'h': []}
    names = {'ws': [], 'bs': [], 'wt': [], 'bt': [], 'h': []}
    l = 0
    while g.has(f'param/source_crossunit_linear.{l}.weight'):
        for key, nm in (('ws', f'source_crossunit_linear.{l}.weight'), ('bs', f'source_crossunit_linear.{l}.bias'),
                        ('wt', f'target_crossunit_linear.{l}.weight'), ('bt', f'target_crossunit_linear.{l}.bias'),
                        ('h', f'crossparas.{l}.weight')):
            p[key].append(g.param(nm))
            names[key].append(nm)
        l += 1
    for key, nm in (('out_s_w', 'source_outputunit.0.weight'), ('out_s_b', 'source_outputunit.0.bias'),
                    ('out_t_w', 'target_outputunit.0.weight'), ('out_t_b', 'target_outputunit.0.bias')):
        p[key] = g.param(nm)
        names[key] = nm
    return p, names


def dtcdr_params(g):
    p, names = {}, {}
    for dom, s in (('source', 's'), ('target', 't')):
        ws, bs, wn, bn = [], [], [], []
        k = 1
        while g.has(f'param/{dom}_mlp_layers.mlp_layers.{k}.weight'):
            wn.append(f'{dom}_mlp_layers.mlp_layers.{k}.weight')
            bn.append(f'{dom}_mlp_layers.mlp_layers.{k}.bias')
            ws.append(g.param(wn[-1]))
            bs.append(g.param(bn[-1]))
            k += 3
        p[f'{s}_mlp_w'], p[f'{s}_mlp_b'] = ws, bs
        names[f'{s}_mlp_w'], names[f'{s}_mlp_b'] = wn, bn
        names[f'{s}_out_w'], names[f'{s}_out_b'] = f'{dom}_predict_layer.weight', f'{dom}_predict_layer.bias'
        p[f'{s}_out_w'], p[f'{s}_out_b'] = g.param(names[f'{s}_out_w']), g.param(names[f'{s}_out_b'])
    return p, names


def bitgcf_graph(g):
    """Degree vectors as bitgcf.py:79-82 builds them ([N,1] fp32) and the per-domain edge lists."""
    n_users = g.param('source_user_embedding.weight').shape[0]
    n_items = g.param('source_item_embedding.weight').shape[0]
    edges, deg = {}, {}
    for dom, s in (('source', 's'), ('target', 't')):
        r, c = g.z[f'edges/{dom}_row'], g.z[f'edges/{dom}_col']
        edges[dom] = (r, c)
        deg[s + 'u'] = torch.from_numpy(np.bincount(r, minlength=n_users).astype(np.float32)).unsqueeze(1)
        deg[s + 'i'] = torch.from_numpy(np.bincount(c, minlength=n_items).astype(np.float32)).unsqueeze(1)
    return n_users, n_items, edges, deg
