# -*- coding: utf-8 -*-
"""
ctypes binding of ``libplsb200.so`` (C ABI declared in include/plsb200.h).

The library is the only compute path of this package: if it cannot be loaded
the import of :mod:`pypyls_b200.engine` fails loudly -- there is no CPU
fallback.
"""

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libplsb200.so')

PLSB_BEHAVIORAL_CORR = 0
PLSB_BEHAVIORAL_COV = 1
PLSB_MEANCENTERED = 2
PLSB_SIMPLS = 3

PLSB_GEMM_AUTO = 0
PLSB_GEMM_DMMA = 1

_vp, _i, _i64, _u64, _dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); mirrors include/plsb200.h one to one
PROTOTYPES = {
    'plsb_version': (_i, []),
    'plsb_last_error': (C.c_char_p, []),
    'plsb_create': (_i, [C.POINTER(_vp), _i]),
    'plsb_destroy': (_i, [_vp]),
    'plsb_set_workspace_limit': (_i, [_vp, _u64]),
    'plsb_set_gemm_backend': (_i, [_vp, _i, _i]),
    'plsb_gemm_work': (_i, [_vp, C.POINTER(_dbl), C.POINTER(_dbl), _i]),
    'plsb_configure': (_i, [_vp, _i, _i, _i, _i, _i, _ip, _i, _i, _i]),
    'plsb_set_data': (_i, [_vp, _vp, _vp, _vp]),
    'plsb_decompose': (_i, [_vp, _vp, _vp, _vp, _vp]),
    'plsb_set_original': (_i, [_vp, _vp, _vp, _vp, _vp]),
    'plsb_project_scores': (_i, [_vp, _vp, _i, _vp, _vp]),
    'plsb_gen_perm_indices': (_i, [_vp, _u64, _i64, _i, _vp, _ip, _vp]),
    'plsb_gen_boot_indices': (_i, [_vp, _u64, _i64, _i, _vp, _ip, _vp]),
    'plsb_run_perms': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'plsb_run_perms_gram': (_i, [_vp, _vp, _i, _vp, _vp]),
    'plsb_run_perms_prepermuted': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'plsb_run_boots': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    'plsb_boot_distrib': (_i, [_vp, _vp, _i, _vp, _vp]),
    'plsb_crossval': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    'plsb_gen_split_masks': (_i, [_vp, _u64, _i64, _i, _i, _dbl, _vp, _ip, _vp]),
    'plsb_split_half': (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    'plsb_boot_chunk': (_i, [_vp, _i]),
    'plsb_perm_pvals': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    'plsb_percentile': (_i, [_vp, _vp, _i, _i, _dbl, _dbl, _vp, _vp, _vp]),
    'plsb_percentile_series': (_i, [_vp, _vp, _i, _i, _i64, _dbl, _dbl, _vp, _vp, _vp]),
    'plsb_transpose': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'plsb_boot_ratio': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    'plsb_dgemm': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    'plsb_gemm_probe': (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, C.POINTER(_dbl), _vp]),
    'plsb_crosscov': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'plsb_small_decomp': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'plsb_gram_proj': (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    'plsb_accum_u': (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    'plsb_timing_enable': (_i, [_vp, _i]),
    'plsb_timing_classes': (_i, []),
    'plsb_timing_class_name': (C.c_char_p, [_i]),
    'plsb_timing_read': (_i, [_vp, C.POINTER(_dbl), C.POINTER(_i64)]),
    'plsb_simpls_set_row_mask': (_i, [_vp, _vp, _vp, _vp]),
    'plsb_simpls_set_original': (_i, [_vp, _vp, _vp]),
    'plsb_simpls_decompose': (_i, [_vp, _vp, _vp, _vp, _vp]),
    'plsb_simpls_run_perms': (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    'plsb_simpls_run_boots': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'plsb_simpls_run_boots_yres': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'plsb_gen_gaussian_tables': (_i, [_vp, _i64, _i, _i, _vp, _vp]),
    'plsb_launch_count': (_i64, [_vp]),
}


class PLSBError(RuntimeError):
    """A call into libplsb200.so returned a non-zero status."""


def load(path=LIB_PATH):
    """Loads the shared library and attaches the prototypes."""
    if not os.path.exists(path):
        raise ImportError(
            'libplsb200.so is missing ({}). Build it with '
            '`python -m pypyls_b200._build` (needs nvcc, sm_100a); there is '
            'no CPU fallback.'.format(path))
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is absent
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def check(status):
    if status != 0:
        msg = lib().plsb_last_error()
        msg = msg.decode('utf-8', 'replace') if msg else ''
        if status == -1:
            raise ValueError(msg)
        raise PLSBError('libplsb200 status {}: {}'.format(status, msg))
