"""merge-spmv_b200: Blackwell-native (sm_100a) merge-based CSR SpMV, a drop-in for the
``cub::DeviceSpmv::CsrMV`` path of dumerrill/merge-spmv.

The product is ``libmergespmv.so`` (C ABI in ``include/mergespmv.h``, kernels in ``csrc/``) and
the C++ drivers in ``host/``.  This Python package is plumbing only: a ctypes binding, the
host-side mirror of the reference operator interface, synthetic matrix generators, and the
one-process-per-GPU sharding built on ``torch.distributed``.

There is no CPU fallback: every compute entry point raises if ``libmergespmv.so`` is missing or
CUDA reports an error.
"""
from . import _lib
from ._lib import MergeSpmvError, lib, lib_path
from .csrmv import DeviceSpmv, MultiGpuSpmvSession, SpmvSession, csrmv, temp_storage, merge_path_search, swath_coords
from . import generators
from . import sharded

__all__ = [
    "DeviceSpmv", "SpmvSession", "MultiGpuSpmvSession", "temp_storage", "csrmv", "merge_path_search", "swath_coords", "generators",
    "sharded", "lib", "lib_path", "MergeSpmvError",
]
