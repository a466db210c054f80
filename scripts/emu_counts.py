"""Instruction / traffic counts of the kernels written without GPU access, taken from the CPU CTA emulator (tests/emu):
warp-level MMA instructions and bytes moved by the row load / RED helpers, per batch row.  A pre-measurement estimate
(no timing): it checks the algorithmic byte counts DESIGN.md quotes and compares the tile engines.

    python scripts/emu_counts.py            # prints a markdown table (profiles/r1_emulator_counts.md is its output)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'recbole-cdr_b200')); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch

import emu_util
import test_emu_conet as TC
import test_emu_mlp as TM

ROWS = []


def add(name, rows, c):
    mma = c['mma_tf32'] + c['mma_bf16']
    ROWS.append(f"| {name} | {rows} | {mma / rows:.1f} | {(c['umma_tf32'] + c['umma_bf16']) / rows:.2f} | {c['row_load_bytes'] / rows:.0f} | "
                f"{c['row_red_bytes'] / rows:.0f} | {c['cta_barriers']} |")


def conet(mode, overlap, batch=256, dim=128, hidden=(64, 32, 16, 8)):
    dims, tabs, P, user, item, label, n_ov, *_ = TC.make_case(batch, dim, list(hidden), 0, True)
    n_ov = {'all': 10 ** 6, 'half': 45, 'none': 0}[overlap]
    Pk = dict(ws=P['ws'], bs=P['bs'], wt=P['wt'], bt=P['bt'], h=P['h'], out_w=P['out_s_w'], out_b=P['out_s_b'])
    Pk = {k: ([x.numpy() for x in v] if isinstance(v, list) else v.numpy()) for k, v in Pk.items()}
    with emu_util.tc_mode(mode):
        emu_util.config(sms=2, seed=0)
        emu_util.counters()
        emu_util.conet_step(dims, Pk, 0, (tabs['source_user'].numpy(), tabs['source_item'].numpy(), tabs['target_user'].numpy(),
                                          tabs['target_item'].numpy()), user.numpy(), item.numpy(), label.numpy(),
                            mask_on_item=False, n_overlap=n_ov)
        return emu_util.counters()


def mlp(mode, which, batch=256):
    with emu_util.tc_mode(mode):
        emu_util.config(sms=2, seed=0)
        if which == 'map':
            src, tgt, ws, bs, idx, leaves, ref = TM.map_case(batch)
            emu_util.counters()
            emu_util.mlp_step(1, [64, 128, 64], [w.numpy() for w in ws], [b.numpy() for b in bs], TM.ACT_TANH, 0, 0,
                              (src.numpy(), None, None, None, tgt.numpy()), idx.numpy(), None, None, tile_rows=64)
        else:
            tabs, ws, bs, u, i, label, *_ = TM.dtcdr_case(batch)
            emu_util.counters()
            emu_util.mlp_step(1, [128, 32, 16, 1], [w.numpy() for w in ws], [b.numpy() for b in bs], TM.ACT_RELU, 1, 1,
                              (tabs['source_user'].numpy(), tabs['target_user'].numpy(), tabs['source_item'].numpy(),
                               tabs['target_item'].numpy(), None), u.numpy(), i.numpy(), label.numpy(), tile_rows=64)
        return emu_util.counters()


def topk(engine, B=256, n_items=2049, D=64, k=20):
    import test_emu_topk as TT
    rng = np.random.RandomState(0)
    U = (rng.randn(B, D) * 0.3).astype(np.float32)
    I = (rng.randn(n_items, D) * 0.3).astype(np.float32)
    emu_util.counters()
    TT.run_emu(U, I, k, sms=2, engine=engine)
    return emu_util.counters()


def map_tc5(batch=256):
    import test_emu_tc5_mlp as T5
    emu_util.counters()
    T5.run(64, batch, 500, seed=1, sms=2)
    return emu_util.counters()


names = {0: '3xTF32 (m16n8k8)', 1: 'bf16x3 (m16n8k16)', 2: '1xTF32 (diagnostic)'}
for mode in (0, 1, 2):
    for ov in ('all', 'half', 'none'):
        add(f'CoNet tower pass fwd+bwd, D = 128, [64,32,16,8], {names[mode]}, overlapped rows: {ov}', 256, conet(mode, ov))
for mode in (0, 1):
    add(f'EMCDR map step fwd+bwd, 64-128-64, {names[mode]}', 256, mlp(mode, 'map'))
    add(f'DTCDR NeuMF term fwd+bwd, 128-32-16-1, {names[mode]}', 256, mlp(mode, 'dtcdr'))
add('EMCDR map step fwd+bwd, 64-128-64, tcgen05 kind::f16 bf16x3 (tc5_mlp.cu, 128-row tiles)', 256, map_tc5())
for eng in ('mma', 'tc5'):
    c = topk(eng)
    add(f'full-sort top-20, 256 users x 2048 items, D = 64, engine {eng} (per user)', 256, c)
print('| kernel | rows | warp-level mma.sync per row | tcgen05.mma per row | row-load bytes per row | RED bytes per row | CTA barriers |')
print('|---|---|---|---|---|---|---|')
print('\n'.join(ROWS))
