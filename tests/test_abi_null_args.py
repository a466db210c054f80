"""Every compute entry point of include/xdr.h validates its arguments on the host BEFORE the first CUDA call: with null
pointers and any size value it must return XDR_ERR_INVALID / XDR_ERR_UNSUPPORTED and a message -- not crash, and not get as
far as the CUDA runtime (which, on this GPU-less box, would answer XDR_ERR_CUDA).  Runs in a child process so that a
segfault shows up as a failed test instead of taking pytest down."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import ctypes, json, sys
sys.path.insert(0, %(pkg)r)
from recbole_cdr_b200 import _lib
QUERIES = {'xdr_version', 'xdr_workspace_bytes', 'xdr_steps_workspace_bytes', 'xdr_topk_workspace_bytes',
           'xdr_tc_conet_scratch_bytes', 'xdr_last_error', 'xdr_device_info', 'xdr_fused_mlp_supported',
           'xdr_tc_mlp_supported', 'xdr_tc_conet_supported', 'xdr_set_coop_launch', 'xdr_set_dense_engine', 'xdr_steps_set_hot_rows'}
out = {}
for name, (res, argt) in sorted(_lib.PROTOTYPES.items()):
    if name in QUERIES or name.endswith('_supported') or name.endswith('_bytes') or res is not _lib.c_int:
        continue     # queries answer with a value, not a status
    for v in (0, 1, 2, 4, 8, 64, -1, 2 ** 31 - 1):
        args = []
        for t in argt:
            if t in (_lib.c_f32, _lib.c_f64):
                args.append(1.0)
            elif t in (_lib.c_int, _lib.c_i64):
                args.append(v)
            elif t in (_lib.c_sz, ctypes.c_uint64, ctypes.c_uint32):
                args.append(max(v, 0))
            else:                      # any pointer type
                args.append(None)
        rc = getattr(_lib._lib, name)(*args)
        out[f'{name}/{v}'] = [int(rc), _lib.last_error() if rc else '']
        print(json.dumps({f'{name}/{v}': out[f'{name}/{v}']}), flush=True)
'''


@pytest.fixture(scope='module')
def results():
    code = CHILD % {'pkg': os.path.join(ROOT, 'recbole-cdr_b200')}
    p = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    got = {}
    for line in p.stdout.splitlines():
        if line.startswith('{'):
            got.update(json.loads(line))
    assert p.returncode == 0, f'child died (rc={p.returncode}) after {list(got)[-1:]}: {p.stderr[-400:]}'
    return got


def test_every_compute_entry_was_exercised(results):
    names = {k.split('/')[0] for k in results}
    assert len(names) >= 30, sorted(names)
    for must in ('xdr_train_steps', 'xdr_bpr_fwd', 'xdr_gather_rows', 'xdr_scatter_add_rows', 'xdr_spmm_csr', 'xdr_dense_fwd',
                 'xdr_fused_mlp_step', 'xdr_tc_mlp_step', 'xdr_tc5_mlp_step', 'xdr_tc_conet_step', 'xdr_sparse_optim_rows', 'xdr_full_sort_topk',
                 'xdr_neg_sample_uniform', 'xdr_train_steps_sharded'):
        assert must in names


def test_null_pointers_are_rejected_before_any_cuda_call(results):
    # sizes of zero may be a documented no-op (return 0); anything else with null buffers must be refused on the host
    bad = {k: v for k, v in results.items() if not k.endswith('/0') and v[0] not in (-1, -3)}
    assert not bad, bad
    zero_bad = {k: v for k, v in results.items() if k.endswith('/0') and v[0] not in (0, -1, -3)}
    assert not zero_bad, zero_bad


def test_refusals_carry_a_message_naming_the_entry_point(results):
    for k, (rc, msg) in results.items():
        if rc < 0:
            assert msg.startswith(k.split('/')[0] + ':'), (k, msg)
