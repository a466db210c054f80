"""Builds and binds the CPU CTA-emulator build of selected libxdr kernels (tests/emu).  TEST INFRASTRUCTURE ONLY: lets the
CPU suite execute kernel *logic* written without GPU access; hardware parity remains the job of the ``gpu`` tests."""
import ctypes
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_DIR = os.path.join(HERE, 'emu')
BUILD_DIR = os.path.join(EMU_DIR, '_build')
CSRC = os.path.join(ROOT, 'recbole-cdr_b200', 'csrc')

# engine mode of the tensor-core tile primitives (tc_tile.cuh XDR_TC_MODE): 0 = 3xTF32 (default), 1 = bf16x3, 2 = one TF32 pass
_mode = 0
_libs = {}


def _lib_path(mode):
    return os.path.join(BUILD_DIR, 'libxdr_emu.so' if mode == 0 else f'libxdr_emu_m{mode}.so')


# translation units of the emulator library: the harness (which #includes the fused row-tile kernels it drives directly)
# plus the kernel files that only need their real entry points
SEPARATE_TUS = ['pair_score.cu', 'gather_scatter.cu', 'dense.cu', 'graph_prop.cu', 'neg_sample.cu', 'topk_score.cu',
                'steps_persistent.cu', 'tc5_mlp.cu', 'tc5_dense.cu']


def _deps():
    return glob.glob(os.path.join(EMU_DIR, '*.cpp')) + glob.glob(os.path.join(EMU_DIR, '*.h')) + \
        glob.glob(os.path.join(CSRC, '*.cu')) + glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(ROOT, 'include', 'xdr.h')]


def _stale(path):
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(mode=None):
    mode = _mode if mode is None else mode
    LIB = _lib_path(mode)
    os.makedirs(BUILD_DIR, exist_ok=True)
    if _stale(LIB):
        import concurrent.futures as cf
        flags = ['-O2', '-std=c++17', '-x', 'c++', '-fPIC', '-DXDR_EMU=1', f'-DXDR_TC_MODE={mode}', '-I', EMU_DIR,
                 '-I', os.path.join(ROOT, 'include'), '-Wno-unknown-pragmas', '-Wno-attributes']
        srcs = [os.path.join(EMU_DIR, 'emu_kernels.cpp')] + [os.path.join(CSRC, f) for f in SEPARATE_TUS]
        objs = [os.path.join(BUILD_DIR, f'm{mode}_' + os.path.basename(f) + '.o') for f in srcs]

        def cc(job):
            src, obj = job
            return subprocess.run(['g++'] + flags + ['-c', src, '-o', obj], capture_output=True, text=True)

        with cf.ThreadPoolExecutor(max_workers=6) as ex:
            for r in ex.map(cc, zip(srcs, objs)):
                if r.returncode != 0:
                    raise RuntimeError('emulator build failed:\n' + r.stdout + r.stderr)
        r = subprocess.run(['g++', '-shared', '-o', LIB] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('emulator link failed:\n' + r.stdout + r.stderr)
    return LIB


def lib():
    if _mode not in _libs:
        L = ctypes.CDLL(build(_mode))
        L.emu_last_error.restype = ctypes.c_char_p
        L.emu_config.argtypes = [ctypes.c_int, ctypes.c_uint64]
        _libs[_mode] = L
    return _libs[_mode]


class tc_mode:
    """Context manager: run the emulated kernels with another engine of the tensor-core tile primitives."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        global _mode
        self.prev, _mode = _mode, self.mode
        return self

    def __exit__(self, *exc):
        global _mode
        _mode = self.prev


def counters(reset=True):
    """Work counters of the emulated launches since the last reset (see cuda_emu.h ``Counters``)."""
    out = (ctypes.c_uint64 * 7)()
    lib().emu_counters(out, ctypes.c_int(1 if reset else 0))
    return dict(zip(('mma_tf32', 'mma_bf16', 'umma_tf32', 'row_load_bytes', 'row_red_bytes', 'cta_barriers', 'umma_bf16'), [int(v) for v in out]))


def config(sms=4, seed=0):
    lib().emu_config(int(sms), int(seed))


def p(a):
    """Host pointer of a numpy array (None -> NULL)."""
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def ptr_array(arrs):
    return (ctypes.c_void_p * len(arrs))(*[None if a is None else a.ctypes.data for a in arrs])


WS_BYTES = 64 + 4 * 4 * 2048


def workspace():
    return np.zeros(WS_BYTES, dtype=np.uint8)


def mlp_step(impl, dims, Ws, bs, hidden_act, in_mode, head, tabs, idx_u, idx_i, label, backward=True, grad_loss=1.0,
             scale=1.0, tile_rows=0):
    """Runs fused_mlp_kernel (impl 0) or tc_mlp_kernel (impl 1) under the emulator.  Returns dict(loss, dW, db, dtabs, prob)."""
    L = lib()
    Au, Bu, Ai, Bi, T = tabs
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    Au, Bu, Ai, Bi, T = map(f32, (Au, Bu, Ai, Bi, T))
    Ws = [f32(w) for w in Ws]
    bs = [f32(b) for b in bs]
    dWs = [np.zeros_like(w) for w in Ws]
    dbs = [None if b is None else np.zeros_like(b) for b in bs]
    dt = [None if t is None else np.zeros_like(t) for t in (Au, Bu, Ai, Bi, T)]
    idx_u = np.ascontiguousarray(idx_u, dtype=np.int64)
    idx_i = None if idx_i is None else np.ascontiguousarray(idx_i, dtype=np.int64)
    label = f32(label)
    B = idx_u.size
    out8 = np.zeros(8, dtype=np.float32)
    prob = np.zeros(B, dtype=np.float32) if head == 1 else None
    g = np.array([grad_loss], dtype=np.float32)
    ws = workspace()
    oob = np.zeros(1, dtype=np.int32)
    darr = (ctypes.c_int * len(dims))(*[int(d) for d in dims])
    L.emu_mlp_step.argtypes = None
    rc = L.emu_mlp_step(
        ctypes.c_int(impl), ctypes.c_int(len(Ws)), darr, ptr_array(Ws), ptr_array(bs), ptr_array(dWs), ptr_array(dbs),
        ctypes.c_int(hidden_act), ctypes.c_int(in_mode), ctypes.c_int(head), p(Au), p(Bu), p(Ai), p(Bi), p(T),
        ctypes.c_int64(Au.shape[0]), ctypes.c_int64(Ai.shape[0] if Ai is not None else 0), ctypes.c_int(Au.shape[1]),
        p(idx_u), p(idx_i), p(label), ctypes.c_int64(B), ctypes.c_int(1 if backward else 0), p(g), ctypes.c_float(scale),
        p(dt[0]), p(dt[1]), p(dt[2]), p(dt[3]), p(dt[4]), p(prob), p(out8), p(ws), p(oob), ctypes.c_int(tile_rows))
    if rc != 0:
        raise RuntimeError(f'emu_mlp_step rc={rc}: {L.emu_last_error().decode()}')
    assert not ws[:64].any(), 'kernel left the workspace ticket dirty'
    return dict(loss=float(out8[0]), dW=dWs, db=dbs, dtabs=dt, prob=prob, oob=int(oob[0]))


def conet_step(dims, P, want, tabs, user, item, label, mask_on_item, n_overlap, backward=True, grad_loss=1.0, scale=1.0):
    """tc_conet_kernel under the emulator.  ``P``: dict of lists ws, bs, wt, bt, h and out_w [1, d], out_b [1] of the wanted
    tower; ``tabs``: (Su, Si, Tu, Ti).  Returns dict(loss, prob, d<name> ...)."""
    L = lib()
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    Su, Si, Tu, Ti = map(f32, tabs)
    nl = len(P['ws'])
    lists = {k: [f32(x) for x in P[k]] for k in ('ws', 'bs', 'wt', 'bt', 'h')}
    grads = {k: [np.zeros_like(x) for x in v] for k, v in lists.items()}
    ow, ob = f32(P['out_w']), f32(P['out_b'])
    dow, dob = np.zeros_like(ow), np.zeros_like(ob)
    dt = [np.zeros_like(t) for t in (Su, Si, Tu, Ti)]
    user = np.ascontiguousarray(user, dtype=np.int64)
    item = np.ascontiguousarray(item, dtype=np.int64)
    label = f32(label)
    B = user.size
    dz1 = np.full((B, 2 * dims[1]), np.nan, dtype=np.float32)
    prob = np.zeros(B, dtype=np.float32)
    out8 = np.zeros(8, dtype=np.float32)
    g = np.array([grad_loss], dtype=np.float32)
    ws, oob = workspace(), np.zeros(1, dtype=np.int32)
    darr = (ctypes.c_int * len(dims))(*[int(d) for d in dims])
    rc = L.emu_conet_step(
        ctypes.c_int(nl), darr, ptr_array(lists['ws']), ptr_array(lists['bs']), ptr_array(lists['wt']), ptr_array(lists['bt']),
        ptr_array(lists['h']), ptr_array(grads['ws']), ptr_array(grads['bs']), ptr_array(grads['wt']), ptr_array(grads['bt']),
        ptr_array(grads['h']), p(ow), p(ob), p(dow), p(dob), ctypes.c_int(want), p(Su), p(Si), p(Tu), p(Ti),
        ctypes.c_int64(Su.shape[0]), ctypes.c_int64(Si.shape[0]), ctypes.c_int(Su.shape[1]), p(user), p(item), p(label),
        ctypes.c_int64(B), ctypes.c_int(1 if mask_on_item else 0), ctypes.c_int64(n_overlap), ctypes.c_int(1 if backward else 0),
        p(g), ctypes.c_float(scale), p(dt[0]), p(dt[1]), p(dt[2]), p(dt[3]), p(dz1), p(prob), p(out8), p(ws), p(oob))
    if rc != 0:
        raise RuntimeError(f'emu_conet_step rc={rc}: {L.emu_last_error().decode()}')
    assert not ws[:64].any(), 'kernel left the workspace ticket dirty'
    return dict(loss=float(out8[0]), prob=prob, grads=grads, dout_w=dow, dout_b=dob, dtabs=dt, oob=int(oob[0]))


# ---------------------------------------------------------------------------------------------------------------------------
# Running the PYTHON side (recbole_cdr_b200.ops autograd functions, the drop-in model classes) on CPU tensors: every libxdr
# entry point that has an ``emu_xdr_*`` twin in the emulator library is redirected to it; anything else raises.
# ---------------------------------------------------------------------------------------------------------------------------
import contextlib


@contextlib.contextmanager
def patched_ops(sms=3, seed=0):
    """Context manager: ``recbole_cdr_b200.ops`` / ``_lib`` call into the emulator with CPU tensors (test use only)."""
    import torch
    from recbole_cdr_b200 import _lib as xl
    from recbole_cdr_b200 import ops
    L = lib()
    config(sms, seed)
    ws_cache = {}

    def emu_fn(name):
        try:
            fn = getattr(L, name)
        except AttributeError:
            raise RuntimeError(f'{name} is not part of the emulator build: this path needs a GPU')
        res, args = xl.PROTOTYPES[name]
        fn.restype, fn.argtypes = res, args
        return fn

    def call(name, *args):
        rc = emu_fn(name)(*args)
        if rc != 0:
            raise xl.XdrError(f'{name} failed under the emulator (status {rc}): {L.emu_last_error().decode()}')

    def req_f32(t, name):
        if t.dtype != torch.float32:
            raise TypeError(f'{name} must be float32, got {t.dtype}')
        if not t.is_contiguous():
            raise ValueError(f'{name} must be contiguous')

    def ids(t, name):
        if t.dtype != torch.int64:
            raise TypeError(f'{name} must be int64 (torch.LongTensor), got {t.dtype}')
        return t.contiguous()

    def workspace(device):
        if 'ws' not in ws_cache:
            ws_cache['ws'] = torch.zeros(WS_BYTES, dtype=torch.uint8)
        return ws_cache['ws']

    class LibProxy:
        """stands in for the ctypes handle ``_lib._lib`` (only the *_supported probes are called through it)"""

        def __getattr__(self, name):
            return emu_fn(name)

    # every loaded module of the package that bound the _lib / ops helpers by name gets the emulator versions
    import sys
    import importlib
    for extra in ('graph', 'sampler', 'sampler.crossdomain_sampler', 'trainer', 'data', 'shard'):
        importlib.import_module('recbole_cdr_b200.' + extra)
    replace = {xl.call: call, xl.cur_stream: (lambda: None), ops._require_cuda_f32: req_f32, ops._ids: ids,
               xl.workspace: workspace, ops._on_device: (lambda t: True)}
    undo = []
    for name, mod in list(sys.modules.items()):
        if not name.startswith('recbole_cdr_b200') or mod is None:
            continue
        for attr, val in list(vars(mod).items()):
            try:
                new = replace.get(val)
            except TypeError:       # unhashable module attribute
                continue
            if new is not None:
                undo.append((mod, attr, val))
                setattr(mod, attr, new)
    handle = xl._lib
    xl._lib = LibProxy()
    try:
        yield ops
    finally:
        for mod, attr, val in undo:
            setattr(mod, attr, val)
        xl._lib = handle
        config(4, 0)
