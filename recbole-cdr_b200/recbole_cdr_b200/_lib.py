"""ctypes binding of libxdr.so -- the only door between Python and the CUDA hot path.

Mirrors ``include/xdr.h`` one to one (same names, same argument order).  Importing this module never needs a
GPU; *calling* any compute entry point does.  If the shared library is missing the import fails loudly -- there
is no Python/PyTorch fallback for the hot path by design.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('XDR_LIB', os.path.join(_HERE, 'lib', 'libxdr.so'))

c_i64, c_int, c_f32, c_vp, c_sz = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
c_f64 = ctypes.c_double

# name -> (restype, argtypes); argument order is exactly that of include/xdr.h
PROTOTYPES = {
    'xdr_version': (c_int, []),
    'xdr_last_error': (ctypes.c_char_p, []),
    'xdr_device_info': (c_int, [ctypes.POINTER(c_int)] * 3),
    'xdr_workspace_bytes': (c_sz, []),
    'xdr_gather_rows': (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]),
    'xdr_scatter_add_rows': (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_f32, c_vp, c_vp]),
    'xdr_gather_max2': (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]),
    'xdr_scatter_max2_bwd': (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp]),
    'xdr_bpr_fwd': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_vp, c_vp, c_vp,
                            c_vp, c_vp, c_vp]),
    'xdr_bpr_bwd': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_vp, c_vp, c_vp,
                            c_vp, c_f32, c_vp, c_vp, c_vp]),
    'xdr_point_fwd': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_int, c_f32, c_vp, c_vp, c_vp,
                              c_vp, c_vp]),
    'xdr_point_bwd': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_int, c_f32, c_vp, c_vp, c_vp,
                              c_f32, c_vp, c_vp, c_vp]),
    'xdr_set_dense_engine': (c_int, [c_int]),
    'xdr_dense_fwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_int, c_vp]),
    'xdr_act_bwd': (c_int, [c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
    'xdr_dense_bwd_input': (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_vp]),
    'xdr_dense_bwd_weight': (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_vp]),
    'xdr_mse_rows_fwd': (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    'xdr_mse_rows_bwd': (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_vp, c_f32, c_vp, c_vp, c_vp]),
    'xdr_bce_logit_fwd': (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    'xdr_bce_logit_bwd': (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    'xdr_frob_sum_fwd': (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    'xdr_frob_sum_bwd': (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    'xdr_set_coop_launch': (c_int, [c_int]),
    'xdr_steps_workspace_bytes': (c_sz, [c_int]),
    'xdr_train_steps': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_int,
                                c_f32, c_f32, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp, c_vp]),
    'xdr_train_steps_host': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_i64, c_int, c_f32, c_f32, c_vp, c_f32, c_vp,
                                     c_vp, c_vp, c_vp, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'xdr_steps_set_hot_rows': (c_int, [c_vp, c_int, c_vp, c_int]),
    'xdr_touch_map_bytes': (c_sz, [c_i64, c_i64]),
    'xdr_train_steps_lazy': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_int,
                                     c_f32, c_f32, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp, c_int, c_vp, c_vp]),
    'xdr_train_steps_sharded': (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_i64,
                                        c_i64, c_int, c_int, c_int, c_f32, c_f32, c_vp, c_f32, c_vp, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp]),
    'xdr_gather_rows_sharded': (c_int, [c_vp, c_int, c_i64, c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp]),
    'xdr_ipc_export': (c_int, [c_vp, c_vp, ctypes.POINTER(c_i64)]),
    'xdr_ipc_open': (c_int, [c_vp, ctypes.POINTER(c_vp)]),
    'xdr_ipc_close': (c_int, [c_vp]),
    'xdr_neg_sample_uniform': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, ctypes.c_uint64,
                                       ctypes.c_uint32, c_int, c_vp, c_vp, c_vp]),
    'xdr_spmm_csr': (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    'xdr_spmm_csr_sharded': (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp]),
    'xdr_prop_elementwise': (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_vp]),
    'xdr_transfer_norm_fwd': (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_int, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp,
                                      c_vp, c_vp, c_i64, c_vp]),
    'xdr_transfer_norm_bwd': (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_int, c_f32,
                                      c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'xdr_fused_mlp_supported': (c_int, [c_int, c_vp]),
    'xdr_fused_mlp_step': (c_int, [c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                   c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp,
                                   c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'xdr_tc_mlp_supported': (c_int, [c_int, c_vp]),
    'xdr_tc_mlp_step': (c_int, [c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp,
                                c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'xdr_tc5_mlp_supported': (c_int, [c_int, c_vp]),
    'xdr_tc5_mlp_step': (c_int, [c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                 c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp,
                                 c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'xdr_tc_conet_supported': (c_int, [c_int, c_vp, c_int]),
    'xdr_tc_conet_scratch_bytes': (c_sz, [c_i64, c_int]),
    'xdr_tc_conet_step': (c_int, [c_int, c_vp] + [c_vp] * 10 + [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp,
                                  c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_int, c_i64, c_int, c_vp, c_f32,
                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'xdr_sparse_optim_rows': (c_int, [c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_i64, c_f32,
                                      c_f32, c_f64, c_f64, c_vp, c_vp]),
    'xdr_topk_workspace_bytes': (c_sz, [c_i64, c_int]),
    'xdr_tc5_selftest': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'xdr_tc5_selftest_bf16': (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    'xdr_full_sort_topk_tc5': (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_i64, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'xdr_full_sort_topk': (c_int, [c_vp, c_i64, c_vp, c_i64, c_int, c_i64, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'xdr_select_dot': (c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_int, c_i64, c_vp, c_vp, c_vp]),
}

LOSS_MSE, LOSS_BCE_SIGMOID, LOSS_NONE = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3
OPT_SGD, OPT_ADAGRAD, OPT_LAZY_ADAM = 0, 1, 2
ACT_BY_NAME = {None: ACT_NONE, 'none': ACT_NONE, 'relu': ACT_RELU, 'tanh': ACT_TANH, 'sigmoid': ACT_SIGMOID}


class XdrError(RuntimeError):
    """Non-zero status from libxdr (the reference signals errors with Python exceptions; so do we)."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'libxdr.so not found at {LIB_PATH}: build it with `python recbole-cdr_b200/build.py` '
            '(needs nvcc; no GPU required). There is no CPU/PyTorch fallback for the hot path.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = _load()


def last_error() -> str:
    return _lib.xdr_last_error().decode()


def version() -> int:
    return _lib.xdr_version()


def workspace_bytes() -> int:
    return _lib.xdr_workspace_bytes()


def call(name: str, *args):
    """Invoke an int-returning entry point and raise XdrError(xdr_last_error()) on failure."""
    rc = getattr(_lib, name)(*args)
    if rc != 0:
        raise XdrError(f'{name} failed (status {rc}): {last_error()}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def cur_stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _device_key(device: torch.device) -> int:
    if device.index is not None:
        return device.index
    return torch.cuda.current_device() if device.type == 'cuda' else -1


_workspaces = {}


def workspace(device: torch.device) -> torch.Tensor:
    """Zero-initialised scratch private to (device, current stream); kernels leave it clean (xdr.h `ws`)."""
    key = (_device_key(device), cur_stream())
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(workspace_bytes(), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


_oob_flags = {}


def oob_flag(device: torch.device) -> torch.Tensor:
    key = _device_key(device)
    f = _oob_flags.get(key)
    if f is None:
        f = torch.zeros(1, dtype=torch.int32, device=device)
        _oob_flags[key] = f
    return f


def check_ids(device: torch.device):
    """Synchronising check of the out-of-range flag; raises IndexError like nn.Embedding does on CPU."""
    f = oob_flag(device)
    if int(f.item()) != 0:
        f.zero_()
        raise IndexError('index out of range in self')
