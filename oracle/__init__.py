"""CPU oracle for the merge-based CsrMV path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package.  The product (``merge_spmv_b200`` and
``libmergespmv.so``) never does; it fails loudly when its CUDA library is missing.

Two checkers live here, both reached through ctypes:

* ``Oracle``    -- ``liboracle.so``, the plain-C restatement in ``merge_oracle.c`` of the
                   reference's ``MergePathSearch`` (cpu_spmv.cpp:223-245), ``OmpMergeCsrmv``
                   (:292-353), ``SpmvGold`` (:257-277) and ``CompareResults``
                   (utils.h:692-742).
* ``Reference`` -- ``_ref/libref_cpu_spmv*.so``, the reference's own ``cpu_spmv.cpp`` compiled
                   from ``/root/reference`` by ``oracle/Makefile`` (``ref_harness.cpp`` only
                   wraps it).  Present wherever the prebuilt ``.so`` travelled; ``None`` otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Run oracle/Makefile (liboracle.so always; _ref only where /root/reference exists)."""
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")) or os.path.exists(
        "/root/reference/cpu_spmv.cpp"
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def _fp(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", _f32p, C.c_float
    if dtype == np.float64:
        return "f64", _f64p, C.c_double
    raise TypeError(f"unsupported value type {dtype}")


def _csr_args(row_offsets, col, val):
    ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
    ci = np.ascontiguousarray(col, dtype=np.int32)
    va = np.ascontiguousarray(val)
    return ro, ci, va


class Oracle:
    """ctypes view of liboracle.so (the C restatement)."""

    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.oracle_merge_path_search.argtypes = [C.c_int, _i32p, C.c_int, C.c_int,
                                               C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_merge_thread_coords.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p]
        for sfx, fpp, fpc in (("f32", _f32p, C.c_float), ("f64", _f64p, C.c_double)):
            getattr(L, f"oracle_merge_csrmv_{sfx}").argtypes = [
                C.c_int, C.c_int, C.c_int, _i32p, _i32p, fpp, fpp, fpp]
            getattr(L, f"oracle_spmv_gold_{sfx}").argtypes = [
                C.c_int, _i32p, _i32p, fpp, fpp, fpp, fpp, fpc, fpc]
            f = getattr(L, f"oracle_compare_results_{sfx}")
            f.argtypes = [fpp, fpp, C.c_longlong]
            f.restype = C.c_longlong
            t = getattr(L, f"oracle_time_merge_csrmv_{sfx}")
            t.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, fpp, fpp, fpp, C.c_int]
            t.restype = C.c_double
        L.oracle_num_procs.restype = C.c_int

    # -- merge path ---------------------------------------------------------------------
    def merge_path_search(self, diagonal, row_offsets, nnz=None):
        ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
        rows = ro.size - 1
        nnz = int(ro[-1]) if nnz is None else nnz
        x, y = C.c_int(), C.c_int()
        # row_end_offsets = row_offsets + 1 (device_spmv.cuh:148, cpu_spmv.cpp:381)
        ends = np.ascontiguousarray(ro[1:]) if rows > 0 else np.zeros(1, np.int32)
        self.lib.oracle_merge_path_search(int(diagonal), ends, rows, nnz, C.byref(x), C.byref(y))
        return x.value, y.value

    def thread_coords(self, num_threads, row_offsets):
        ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
        rows, nnz = ro.size - 1, int(ro[-1])
        ends = np.ascontiguousarray(ro[1:]) if rows > 0 else np.zeros(1, np.int32)
        out = np.zeros(2 * (num_threads + 1), np.int32)
        self.lib.oracle_merge_thread_coords(num_threads, rows, nnz, ends, out)
        return out.reshape(-1, 2)

    # -- SpMV ---------------------------------------------------------------------------
    def merge_csrmv(self, row_offsets, col, val, x, num_threads=1):
        ro, ci, va = _csr_args(row_offsets, col, val)
        sfx, _, _ = _fp(va.dtype)
        rows, nnz = ro.size - 1, int(ro[-1])
        xv = np.ascontiguousarray(x, dtype=va.dtype)
        y = np.full(max(rows, 1), np.nan, dtype=va.dtype)
        ends = np.ascontiguousarray(ro[1:]) if rows > 0 else np.zeros(1, np.int32)
        if ci.size == 0:
            ci, va = np.zeros(1, np.int32), np.zeros(1, va.dtype)
        getattr(self.lib, f"oracle_merge_csrmv_{sfx}")(num_threads, rows, nnz, ends, ci, va, xv, y)
        return y[:rows]

    def spmv_gold(self, row_offsets, col, val, x, y_in=None, alpha=1.0, beta=0.0):
        ro, ci, va = _csr_args(row_offsets, col, val)
        sfx, _, _ = _fp(va.dtype)
        rows = ro.size - 1
        xv = np.ascontiguousarray(x, dtype=va.dtype)
        yi = np.zeros(max(rows, 1), va.dtype) if y_in is None else np.ascontiguousarray(y_in, va.dtype)
        yo = np.zeros(max(rows, 1), va.dtype)
        if ci.size == 0:
            ci, va = np.zeros(1, np.int32), np.zeros(1, va.dtype)
        getattr(self.lib, f"oracle_spmv_gold_{sfx}")(rows, ro, ci, va, xv, yi, yo, alpha, beta)
        return yo[:rows]

    def compare_results(self, computed, reference):
        """0 = PASS under the reference's tolerance rule, else 1 + index of first failure."""
        a = np.ascontiguousarray(computed)
        sfx, _, _ = _fp(a.dtype)
        b = np.ascontiguousarray(reference, dtype=a.dtype)
        return int(getattr(self.lib, f"oracle_compare_results_{sfx}")(a, b, a.size))

    def time_merge_csrmv(self, row_offsets, col, val, x, num_threads, iterations):
        ro, ci, va = _csr_args(row_offsets, col, val)
        sfx, _, _ = _fp(va.dtype)
        rows, nnz = ro.size - 1, int(ro[-1])
        xv = np.ascontiguousarray(x, dtype=va.dtype)
        y = np.zeros(rows, va.dtype)
        ms = getattr(self.lib, f"oracle_time_merge_csrmv_{sfx}")(
            num_threads, rows, nnz, np.ascontiguousarray(ro[1:]), ci, va, xv, y, iterations)
        return float(ms), y

    def num_procs(self):
        return int(self.lib.oracle_num_procs())


def _host_has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line.split()
                    return "avx2" in flags and "fma" in flags and "bmi2" in flags
    except OSError:
        pass
    return False


class Reference:
    """ctypes view of oracle/_ref/libref_cpu_spmv*.so (the reference's own compiled code)."""

    KINDS = {"market": 0, "grid2d": 1, "grid3d": 2, "wheel": 3, "dense": 4}

    @staticmethod
    def path():
        names = ["libref_cpu_spmv.so"]
        if _host_has_avx2():
            names.insert(0, "libref_cpu_spmv_v3.so")
        for n in names:
            p = os.path.join(_HERE, "_ref", n)
            if os.path.exists(p):
                return p
        return None

    @classmethod
    def available(cls) -> bool:
        return cls.path() is not None

    def __init__(self):
        p = self.path()
        if p is None:
            raise FileNotFoundError("oracle/_ref is not built (needs /root/reference; run make -C oracle)")
        self.so_path = p
        self.lib = C.CDLL(p)
        L = self.lib
        L.ref_merge_path_search.argtypes = [C.c_int, _i32p, C.c_int, C.c_int, _i32p]
        for sfx, fpp, fpc in (("f32", _f32p, C.c_float), ("f64", _f64p, C.c_double)):
            getattr(L, f"ref_omp_merge_csrmv_{sfx}").argtypes = [
                C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, fpp, fpp, fpp]
            getattr(L, f"ref_spmv_gold_{sfx}").argtypes = [
                C.c_int, C.c_int, C.c_int, _i32p, _i32p, fpp, fpp, fpp, fpp, fpc, fpc]
            t = getattr(L, f"ref_time_omp_merge_csrmv_{sfx}")
            t.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, fpp, fpp, fpp, fpp, C.c_int]
            t.restype = C.c_double
            getattr(L, f"ref_compare_results_{sfx}").argtypes = [fpp, fpp, C.c_int]
        L.ref_build.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, _i32p]
        L.ref_built_copy.argtypes = [_i32p, _i32p, _f64p]
        L.ref_built_stats.argtypes = [_f64p]

    def merge_path_search(self, diagonal, row_offsets):
        ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
        rows, nnz = ro.size - 1, int(ro[-1])
        ends = np.ascontiguousarray(ro[1:]) if rows > 0 else np.zeros(1, np.int32)
        out = np.zeros(2, np.int32)
        self.lib.ref_merge_path_search(int(diagonal), ends, rows, nnz, out)
        return int(out[0]), int(out[1])

    def omp_merge_csrmv(self, row_offsets, col, val, x, num_threads=1, num_cols=None):
        ro, ci, va = _csr_args(row_offsets, col, val)
        sfx, _, _ = _fp(va.dtype)
        rows, nnz = ro.size - 1, int(ro[-1])
        xv = np.ascontiguousarray(x, dtype=va.dtype)
        y = np.full(max(rows, 1), np.nan, dtype=va.dtype)
        if ci.size == 0:
            ci, va = np.zeros(1, np.int32), np.zeros(1, va.dtype)
        assert num_threads <= 256, "reference carry arrays hold 256 threads (cpu_spmv.cpp:302-303)"
        getattr(self.lib, f"ref_omp_merge_csrmv_{sfx}")(
            num_threads, rows, num_cols or xv.size, nnz, ro.copy(), ci, va, xv, y)
        return y[:rows]

    def spmv_gold(self, row_offsets, col, val, x, y_in=None, alpha=1.0, beta=0.0):
        ro, ci, va = _csr_args(row_offsets, col, val)
        sfx, _, _ = _fp(va.dtype)
        rows, nnz = ro.size - 1, int(ro[-1])
        xv = np.ascontiguousarray(x, dtype=va.dtype)
        yi = np.zeros(max(rows, 1), va.dtype) if y_in is None else np.ascontiguousarray(y_in, va.dtype)
        yo = np.zeros(max(rows, 1), va.dtype)
        if ci.size == 0:
            ci, va = np.zeros(1, np.int32), np.zeros(1, va.dtype)
        getattr(self.lib, f"ref_spmv_gold_{sfx}")(rows, xv.size, nnz, ro, ci, va, xv, yi, yo, alpha, beta)
        return yo[:rows]

    def compare_results(self, computed, reference):
        a = np.ascontiguousarray(computed)
        sfx, _, _ = _fp(a.dtype)
        b = np.ascontiguousarray(reference, dtype=a.dtype)
        return int(getattr(self.lib, f"ref_compare_results_{sfx}")(a, b, a.size))

    def time_omp_merge_csrmv(self, row_offsets, col, val, x, num_threads, iterations):
        ro, ci, va = _csr_args(row_offsets, col, val)
        sfx, _, _ = _fp(va.dtype)
        rows, nnz = ro.size - 1, int(ro[-1])
        xv = np.ascontiguousarray(x, dtype=va.dtype)
        y = np.zeros(rows, va.dtype)
        yref = np.zeros(rows, va.dtype)
        ms = getattr(self.lib, f"ref_time_omp_merge_csrmv_{sfx}")(
            min(num_threads, 256), rows, xv.size, nnz, ro, ci, va, xv, yref, y, iterations)
        return float(ms), y

    def num_procs(self):
        return int(self.lib.ref_num_procs())

    def build_matrix(self, kind, a=0, b=0, path=""):
        """CSR (fp64 values) built by the reference's own CooMatrix/CsrMatrix code."""
        dims = np.zeros(3, np.int32)
        rc = self.lib.ref_build(self.KINDS[kind], path.encode(), int(a), int(b), dims)
        if rc:
            raise ValueError(kind)
        rows, cols, nnz = map(int, dims)
        ro = np.zeros(rows + 1, np.int32)
        ci = np.zeros(max(nnz, 1), np.int32)
        va = np.zeros(max(nnz, 1), np.float64)
        self.lib.ref_built_copy(ro, ci, va)
        stats = np.zeros(4, np.float64)
        self.lib.ref_built_stats(stats)
        self.lib.ref_built_free()
        return dict(rows=rows, cols=cols, nnz=nnz, row_offsets=ro, col=ci[:nnz], val=va[:nnz],
                    stats=stats)
