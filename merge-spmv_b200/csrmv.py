"""Host-side mirror of the reference operator interface for the CsrMV path.

``DeviceSpmv.CsrMV`` has the argument list of ``cub::DeviceSpmv::CsrMV``
(cub/device/device_spmv.cuh:129-164) and the same two-phase temp-storage protocol as used by
``TestGpuMergeCsrmv`` (gpu_spmv.cu:390-429); tensors stand in for raw device pointers.  All
work happens in ``libmergespmv.so`` through the C ABI -- torch only owns memory and streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

cudaSuccess = 0
cudaErrorInvalidValue = 1


def _ptr(t):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError("expected a torch.Tensor")
    if not t.is_cuda:
        raise _lib.MergeSpmvError(
            "merge_spmv_b200 computes on CUDA devices only (no CPU fallback); got a CPU tensor")
    if not t.is_contiguous():
        raise ValueError("tensors must be contiguous")
    return C.c_void_p(t.data_ptr())


def _stream(stream):
    if stream is None:
        stream = torch.cuda.current_stream()
    if isinstance(stream, torch.cuda.Stream):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


def _suffix(values_dtype):
    if values_dtype == torch.float32:
        return "f32"
    if values_dtype == torch.float64:
        return "f64"
    raise TypeError(f"ValueT must be float32 or float64 (gpu_spmv.cu:730,734), got {values_dtype}")


class DeviceSpmv:
    """Drop-in for ``cub::DeviceSpmv`` (device_spmv.cuh:70-167): only ``CsrMV`` exists."""

    @staticmethod
    def CsrMV(d_temp_storage, temp_storage_bytes, d_values, d_row_offsets, d_column_indices,
              d_vector_x, d_vector_y, num_rows, num_cols, num_nonzeros, stream=None,
              debug_synchronous=False, *, dtype=None):
        """Returns ``(cudaError, temp_storage_bytes)``.

        ``d_temp_storage is None`` -> size query, nothing is launched (dispatch_spmv_orig.cuh:651-655;
        ``dtype`` names ValueT when no value tensor is given).  Otherwise runs ``y = A*x`` on
        ``stream`` (default: torch's current stream), asynchronously unless ``debug_synchronous``.
        """
        vt = d_values.dtype if d_values is not None else (dtype or torch.float64)
        fn = getattr(_lib.lib(), f"mspmv_csrmv_{_suffix(vt)}")
        nbytes = C.c_size_t(int(temp_storage_bytes))
        if d_temp_storage is None:
            err = fn(None, C.byref(nbytes), None, None, None, None, None, int(num_rows), int(num_cols),
                     int(num_nonzeros), None, 0)
            return err, nbytes.value
        if d_vector_x.dtype != vt or d_vector_y.dtype != vt:
            raise TypeError("values, x and y must share ValueT")
        if d_row_offsets.dtype != torch.int32 or d_column_indices.dtype != torch.int32:
            raise TypeError("OffsetT is int32 (gpu_spmv.cu:730,734)")
        with torch.cuda.device(d_vector_y.device):
            err = fn(_ptr(d_temp_storage), C.byref(nbytes), _ptr(d_values), _ptr(d_row_offsets),
                     _ptr(d_column_indices), _ptr(d_vector_x), _ptr(d_vector_y), int(num_rows),
                     int(num_cols), int(num_nonzeros), _stream(stream), int(bool(debug_synchronous)))
        return err, nbytes.value

    @staticmethod
    def CsrMVAxpby(d_temp_storage, temp_storage_bytes, d_values, d_row_offsets, d_column_indices,
                   d_vector_x, d_vector_y, num_rows, num_cols, num_nonzeros, alpha, beta,
                   stream=None, debug_synchronous=False, *, dtype=None):
        """``y = alpha*A*x + beta*y`` (the --alpha/--beta surface, gpu_spmv.cu:721-722, :72-92)."""
        vt = d_values.dtype if d_values is not None else (dtype or torch.float64)
        fn = getattr(_lib.lib(), f"mspmv_csrmv_axpby_{_suffix(vt)}")
        nbytes = C.c_size_t(int(temp_storage_bytes))
        if d_temp_storage is None:
            err = fn(None, C.byref(nbytes), None, None, None, None, None, int(num_rows), int(num_cols),
                     int(num_nonzeros), float(alpha), float(beta), None, 0)
            return err, nbytes.value
        with torch.cuda.device(d_vector_y.device):
            err = fn(_ptr(d_temp_storage), C.byref(nbytes), _ptr(d_values), _ptr(d_row_offsets),
                     _ptr(d_column_indices), _ptr(d_vector_x), _ptr(d_vector_y), int(num_rows),
                     int(num_cols), int(num_nonzeros), float(alpha), float(beta), _stream(stream),
                     int(bool(debug_synchronous)))
        return err, nbytes.value


_temp_cache = {}
_temp_retired = []  # superseded blobs stay alive: a captured CUDA graph or an in-flight call may still use them


def _temp_for(device, nbytes):
    key = (device.index if device.index is not None else torch.cuda.current_device())
    t = _temp_cache.get(key)
    if t is None or t.numel() < nbytes:
        if t is not None:
            _temp_retired.append(t)
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _temp_cache[key] = t
    return t


def temp_storage(values_dtype, num_rows, num_cols, num_nonzeros, device, axpby=False):
    """A temp blob of the size ``mspmv_csrmv_*`` asks for this shape (the caller owns it)."""
    if axpby:
        err, nbytes = DeviceSpmv.CsrMVAxpby(None, 0, None, None, None, None, None, num_rows, num_cols, num_nonzeros,
                                            1.0, 0.0, dtype=values_dtype)
    else:
        err, nbytes = DeviceSpmv.CsrMV(None, 0, None, None, None, None, None, num_rows, num_cols, num_nonzeros,
                                       dtype=values_dtype)
    _lib.check(err, "CsrMV size query")
    return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)


def csrmv(row_offsets, column_indices, values, x, y=None, *, alpha=None, beta=None, num_cols=None,
          stream=None, debug_synchronous=False, temp=None):
    """Convenience wrapper: size query + temp blob + run.  Returns ``y``.

    ``temp``: the caller's own blob (``temp_storage()``); anything that outlives the call -- a captured
    CUDA graph, work on several streams -- must pass one.  Without it a blob cached per device is used:
    it is shared by every such call on that device, so those calls must be stream-ordered with respect
    to each other (like any CUB temp storage)."""
    rows = row_offsets.numel() - 1
    nnz = values.numel()
    cols = int(num_cols) if num_cols is not None else x.numel()
    if y is None:
        y = torch.empty(rows, dtype=values.dtype, device=values.device)
    axpby = alpha is not None or beta is not None
    if axpby:
        a = 1.0 if alpha is None else alpha
        b = 0.0 if beta is None else beta
        err, nbytes = DeviceSpmv.CsrMVAxpby(None, 0, None, None, None, None, None, rows, cols, nnz, a, b,
                                            dtype=values.dtype)
        _lib.check(err, "CsrMV size query")
        if temp is None:
            temp = _temp_for(values.device, nbytes)
        err, _ = DeviceSpmv.CsrMVAxpby(temp, temp.numel(), values, row_offsets, column_indices, x, y,
                                       rows, cols, nnz, a, b, stream, debug_synchronous)
    else:
        err, nbytes = DeviceSpmv.CsrMV(None, 0, None, None, None, None, None, rows, cols, nnz,
                                       dtype=values.dtype)
        _lib.check(err, "CsrMV size query")
        if temp is None:
            temp = _temp_for(values.device, nbytes)
        err, _ = DeviceSpmv.CsrMV(temp, temp.numel(), values, row_offsets, column_indices, x, y, rows,
                                  cols, nnz, stream, debug_synchronous)
    _lib.check(err, "CsrMV")
    return y


def merge_path_search(row_offsets, diagonals, stream=None):
    """Device MergePathSearch (cpu_spmv.cpp:223-245) for a tensor of diagonals -> (n, 2) int32."""
    rows = row_offsets.numel() - 1
    nnz = int(row_offsets[-1].item()) if rows >= 0 else 0
    d = diagonals.to(device=row_offsets.device, dtype=torch.int32).contiguous()
    out = torch.empty((d.numel(), 2), dtype=torch.int32, device=row_offsets.device)
    with torch.cuda.device(row_offsets.device):
        err = _lib.lib().mspmv_merge_path_search(_ptr(row_offsets), rows, nnz, _ptr(d), d.numel(),
                                                 _ptr(out), _stream(stream))
    _lib.check(err, "merge_path_search")
    return out


def swath_coords(row_offsets, value_bytes, stream=None):
    """Start coordinate of every threadblock swath CsrMV uses for this shape (+ final end)."""
    rows = row_offsets.numel() - 1
    nnz = int(row_offsets[-1].item())
    n = C.c_int(0)
    with torch.cuda.device(row_offsets.device):
        L = _lib.lib()
        _lib.check(L.mspmv_csrmv_swath_coords(None, rows, nnz, value_bytes, C.byref(n), None, None))
        out = torch.empty((n.value + 1, 2), dtype=torch.int32, device=row_offsets.device)
        _lib.check(L.mspmv_csrmv_swath_coords(_ptr(row_offsets), rows, nnz, value_bytes, C.byref(n),
                                              _ptr(out), _stream(stream)))
    return out


def csrmv_config(value_bytes, num_rows, num_nonzeros):
    out = (C.c_int * 6)()
    _lib.check(_lib.lib().mspmv_csrmv_config(value_bytes, num_rows, num_nonzeros, out))
    return dict(blocks=out[0], threads=out[1], tile_items=out[2], smem_bytes=out[3],
                kernels_per_call=out[4], tiles=out[5])


class SpmvSession:
    """Host-buffer operator: upload the CSR once, apply it to host vectors (gpu_spmv.cu:542-556,
    :421-432 behind one handle).  Host arrays are numpy; pinned torch tensors avoid staging."""

    def __init__(self, row_offsets, column_indices, values, num_cols, device=0):
        ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
        ci = np.ascontiguousarray(column_indices, dtype=np.int32)
        va = np.ascontiguousarray(values)
        if va.dtype not in (np.float32, np.float64):
            raise TypeError("ValueT must be float32 or float64")
        self.dtype = va.dtype
        self.rows, self.cols, self.nnz = ro.size - 1, int(num_cols), va.size
        self._h = C.c_void_p()
        _lib.check(_lib.lib().mspmv_session_create(
            C.byref(self._h), device, va.dtype.itemsize, self.rows, self.cols, self.nnz,
            ro.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p),
            va.ctypes.data_as(C.c_void_p)), "session_create")

    @staticmethod
    def _host_ptr(a):
        if isinstance(a, torch.Tensor):
            assert not a.is_cuda and a.is_contiguous()
            return C.c_void_p(a.data_ptr())
        return a.ctypes.data_as(C.c_void_p)

    def apply(self, x_host, y_host):
        _lib.check(_lib.lib().mspmv_session_apply(self._h, self._host_ptr(x_host),
                                                  self._host_ptr(y_host)), "session_apply")
        return y_host

    def apply_many(self, n, xs_host, ys_host):
        _lib.check(_lib.lib().mspmv_session_apply_many(self._h, int(n), self._host_ptr(xs_host),
                                                       self._host_ptr(ys_host)), "session_apply_many")
        return ys_host

    def close(self):
        if self._h:
            _lib.lib().mspmv_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGpuSpmvSession:
    """Host-buffer operator over several GPUs of ONE process (``mspmv_mg_session_*``, csrc/mg_session.cu):
    merge-path shards, peer copies of x, one NVLink carry-exchange kernel per device -- no NCCL, no
    torch.distributed.  ``devices``: the device of every shard (a device may appear more than once)."""

    def __init__(self, row_offsets, column_indices, values, num_cols, devices):
        ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
        ci = np.ascontiguousarray(column_indices, dtype=np.int32)
        va = np.ascontiguousarray(values)
        if va.dtype not in (np.float32, np.float64):
            raise TypeError("ValueT must be float32 or float64")
        self.dtype = va.dtype
        self.rows, self.cols, self.nnz = ro.size - 1, int(num_cols), va.size
        self.devices = [int(d) for d in devices]
        ids = (C.c_int * len(self.devices))(*self.devices)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().mspmv_mg_session_create(
            C.byref(self._h), len(self.devices), ids, va.dtype.itemsize, self.rows, self.cols, self.nnz,
            ro.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), va.ctypes.data_as(C.c_void_p)),
            "mg_session_create")

    _host_ptr = staticmethod(SpmvSession._host_ptr)

    def apply(self, x_host, y_host):
        _lib.check(_lib.lib().mspmv_mg_session_apply(self._h, self._host_ptr(x_host), self._host_ptr(y_host)),
                   "mg_session_apply")
        return y_host

    def apply_many(self, n, xs_host, ys_host):
        _lib.check(_lib.lib().mspmv_mg_session_apply_many(self._h, int(n), self._host_ptr(xs_host),
                                                          self._host_ptr(ys_host)), "mg_session_apply_many")
        return ys_host

    def time_device(self, iterations):
        ms = C.c_float()
        _lib.check(_lib.lib().mspmv_mg_session_time_device(self._h, int(iterations), C.byref(ms)), "mg_session_time_device")
        return float(ms.value)

    def shard(self, g):
        out = (C.c_int * 5)()
        _lib.check(_lib.lib().mspmv_mg_session_shard(self._h, int(g), out), "mg_session_shard")
        return dict(x0=out[0], y0=out[1], x1=out[2], y1=out[3], device=out[4])

    def close(self):
        if self._h:
            _lib.lib().mspmv_mg_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
