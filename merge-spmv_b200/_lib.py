"""ctypes binding of libmergespmv.so -- every symbol include/mergespmv.h declares."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (restype, argtypes); the single source of truth the "exports every symbol" test checks
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
_csrmv = [_vp, C.POINTER(_sz), _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i]
SIGNATURES = {
    "mspmv_csrmv_f32": (_i, _csrmv),
    "mspmv_csrmv_f64": (_i, _csrmv),
    "mspmv_csrmv_axpby_f32": (_i, _csrmv[:10] + [C.c_float, C.c_float, _vp, _i]),
    "mspmv_csrmv_axpby_f64": (_i, _csrmv[:10] + [C.c_double, C.c_double, _vp, _i]),
    "mspmv_merge_path_search": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp]),
    "mspmv_csrmv_swath_coords": (_i, [_vp, _i, _i, _i, C.POINTER(_i), _vp, _vp]),
    "mspmv_host_merge_path_search": (None, [_vp, _i, _i, C.c_int64, C.POINTER(_i), C.POINTER(_i)]),
    "mspmv_shard_partition": (None, [_vp, _i, _i, _i, _vp]),
    "mspmv_shard_row_offsets": (None, [_vp, _i, _i, _i, _i, _vp]),
    "mspmv_apply_carries_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "mspmv_apply_carries_f64": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "mspmv_exchange_buffer_bytes": (_sz, [_i]),
    "mspmv_exchange_carries_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "mspmv_exchange_carries_f64": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "mspmv_session_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "mspmv_session_apply": (_i, [_vp, _vp, _vp]),
    "mspmv_session_apply_many": (_i, [_vp, _i, _vp, _vp]),
    "mspmv_session_destroy": (None, [_vp]),
    "mspmv_mg_session_create": (_i, [C.POINTER(_vp), _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "mspmv_mg_session_apply": (_i, [_vp, _vp, _vp]),
    "mspmv_mg_session_apply_many": (_i, [_vp, _i, _vp, _vp]),
    "mspmv_mg_session_time_device": (_i, [_vp, _i, C.POINTER(C.c_float)]),
    "mspmv_mg_session_shard": (_i, [_vp, _i, _vp]),
    "mspmv_mg_session_destroy": (None, [_vp]),
    "mspmv_host_alloc": (_i, [C.POINTER(_vp), _sz]),
    "mspmv_host_free": (_i, [_vp]),
    "mspmv_version": (_i, []),
    "mspmv_ptx_version": (_i, [C.POINTER(_i)]),
    "mspmv_launch_count": (C.c_uint64, []),
    "mspmv_csrmv_config": (_i, [_i, _i, _i, C.POINTER(_i)]),
    "mspmv_error_string": (C.c_char_p, [_i]),
    "mspmv_set_engine": (_i, [C.c_char_p]),
    "mspmv_set_option": (_i, [C.c_char_p, _i]),
}


class MergeSpmvError(RuntimeError):
    pass


def lib_path() -> str:
    # MSPMV_LIB picks an alternative build of the same library (tuning experiments only)
    return os.environ.get("MSPMV_LIB") or os.path.join(_HERE, "libmergespmv.so")


_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise MergeSpmvError(
                f"{p} is missing: build it with `make -C merge-spmv_b200` (or __graft_entry__.build()). "
                "There is no CPU fallback."
            )
        L = C.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(err: int, what: str = "") -> None:
    if err != 0:
        msg = lib().mspmv_error_string(err).decode()
        raise MergeSpmvError(f"{what or 'libmergespmv'} failed: cudaError {err} ({msg})")
