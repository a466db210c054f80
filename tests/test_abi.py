"""CPU-side checks of the C-ABI boundary: libxdr.so loads, exports every symbol include/xdr.h declares, and the
ctypes prototypes of the Python binding agree with the header (no compute calls -- no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'xdr.h')


def declared():
    """[(name, n_args)] parsed from the XDR_API prototypes of include/xdr.h."""
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    out = []
    for m in re.finditer(r'XDR_API\s+[\w\s\*]+?\b(xdr_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ('', 'void') else args.count(',') + 1
        out.append((m.group(1), n))
    return out


def test_header_declares_something():
    names = [n for n, _ in declared()]
    assert 'xdr_bpr_fwd' in names and 'xdr_gather_rows' in names and len(names) >= 20


def test_library_exports_every_declared_symbol():
    from recbole_cdr_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name, _ in declared():
        assert hasattr(lib, name), f'{name} declared in xdr.h but not exported by libxdr.so'


def test_python_prototypes_match_header():
    from recbole_cdr_b200 import _lib
    decl = dict(declared())
    assert set(decl) == set(_lib.PROTOTYPES), set(decl) ^ set(_lib.PROTOTYPES)
    for name, (_, argtypes) in _lib.PROTOTYPES.items():
        assert len(argtypes) == decl[name], f'{name}: header has {decl[name]} args, binding has {len(argtypes)}'


def declared_types():
    """{name: [kind per argument]} with kind in int / i64 / f32 / size / u64 / u32 / ptr, parsed from include/xdr.h."""
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'XDR_API\s+[\w\s\*]+?\b(xdr_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
        args = m.group(2).strip()
        kinds = []
        if args not in ('', 'void'):
            for a in args.split(','):
                a = ' '.join(a.split())
                if '*' in a or 'xdr_stream_t' in a:
                    kinds.append('ptr')
                elif 'int64_t' in a and 'uint' not in a:
                    kinds.append('i64')
                elif 'uint64_t' in a:
                    kinds.append('u64')
                elif 'uint32_t' in a:
                    kinds.append('u32')
                elif 'size_t' in a:
                    kinds.append('size')
                elif 'double' in a:
                    kinds.append('f64')
                elif 'float' in a:
                    kinds.append('f32')
                elif re.match(r'(const )?int\b', a):
                    kinds.append('int')
                else:
                    raise AssertionError(f'{m.group(1)}: unparsed argument {a!r}')
        out[m.group(1)] = kinds
    return out


def test_python_prototype_types_match_header():
    """Same width and class per argument: a c_int where the header says int64_t would corrupt the call silently."""
    from recbole_cdr_b200 import _lib
    kind_of = {ctypes.c_int: 'int', ctypes.c_int64: 'i64', ctypes.c_float: 'f32', ctypes.c_size_t: 'size',
               ctypes.c_uint64: 'u64', ctypes.c_uint32: 'u32', ctypes.c_double: 'f64', ctypes.c_void_p: 'ptr', ctypes.c_char_p: 'ptr'}
    decl = declared_types()
    for name, (_, argtypes) in _lib.PROTOTYPES.items():
        got = ['ptr' if t not in kind_of else kind_of[t] for t in argtypes]
        # size_t and uint64_t are the same class on this ABI
        norm = lambda ks: ['u64' if k == 'size' else k for k in ks]
        assert norm(got) == norm(decl[name]), f'{name}: header {decl[name]} vs binding {got}'


def test_version_and_workspace():
    from recbole_cdr_b200 import _lib
    assert _lib.version() == 100
    assert _lib.workspace_bytes() >= 64


def test_invalid_arguments_are_rejected_without_a_gpu():
    """Argument validation happens on the host before any CUDA call, so it is checkable here."""
    from recbole_cdr_b200 import _lib
    with pytest.raises(_lib.XdrError, match='dim'):
        _lib.call('xdr_gather_rows', None, 10, 6, None, 1, None, 6, None, None)  # dim % 4 != 0
    with pytest.raises(_lib.XdrError, match='null'):
        _lib.call('xdr_gather_rows', None, 10, 64, None, 1, None, 64, None, None)
    with pytest.raises(_lib.XdrError, match='batch'):
        _lib.call('xdr_bpr_fwd', None, None, 1, 1, 64, None, None, None, 0, 1e-10, 0.0, None, None, None, None, None, None)
    assert 'batch' in _lib.last_error()


def test_ops_refuse_cpu_tensors():
    """The product path has no CPU fallback: CPU tensors are an error, not a slow path."""
    import torch
    from recbole_cdr_b200 import ops
    t = torch.zeros(8, 64)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.gather_rows_raw(t, torch.zeros(2, dtype=torch.int64))


def test_get_model_naming_rule():
    from recbole_cdr_b200.utils import get_model
    for name in ('EMCDR', 'CMF', 'CoNet', 'DTCDR'):
        assert get_model(name).__name__ == name
    with pytest.raises(ValueError):
        get_model('NoSuchModel')


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps each exported entry point to the reference interface it replaces (or says it has none)."""
    from recbole_cdr_b200 import _lib
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    missing = [n for n in _lib.PROTOTYPES if n not in doc]
    assert not missing, missing
