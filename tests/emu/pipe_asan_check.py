"""Run the pipe-engine kernel under the SIMT interpreter with AddressSanitizer (spawned by
tests/test_pipe_emu.py with LD_PRELOAD=libasan.so): every caller array is a fresh, exact-size
allocation with red zones behind its last element, so a kernel that reads or writes outside
row_offsets[0..rows], column_indices / values[0..nnz), x[0..cols), y[0..rows) or outside its own
temporaries aborts the process.  The reference's GPU path has exactly such accesses at the last
row; this is the check that ours does not."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from conftest import random_csr  # noqa: E402

lib = C.CDLL(sys.argv[1])
for sfx, fp in (("f64", C.c_double), ("f32", C.c_float)):
    f = getattr(lib, "emu_pipe_" + sfx)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, fp, fp, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]


def call(v, r, c, x, y, rows, nnz, dt, blocks, search):
    stats = np.zeros(4, np.int32)
    fn = lib.emu_pipe_f64 if dt == np.float64 else lib.emu_pipe_f32
    return fn(v, r, c, x, y.ctypes.data, rows, nnz, 1.0, 0.0, 0, blocks, search, None, stats.ctypes.data)


def strict(ro, col, val, x, blocks, search):
    ro, col = np.array(ro, np.int32, copy=True), np.array(col, np.int32, copy=True)
    val, x = np.array(val, copy=True), np.array(x, copy=True)
    rows, nnz = ro.size - 1, int(ro[-1])
    y = np.full(rows, np.nan, val.dtype)
    assert call(val.ctypes.data, ro.ctypes.data, col.ctypes.data, x.ctypes.data, y, rows, nnz, val.dtype, blocks, search) == 0
    return y

# Misaligned bases (slices of larger arrays, as the multi-GPU shards produce): the slack around the
# view is poisoned by hand, so the aligned-superset staging may not step outside the view either.
asan = C.CDLL(None)
poison = getattr(asan, "__asan_poison_memory_region")      # (looked up by string: inside a class body the
unpoison = getattr(asan, "__asan_unpoison_memory_region")  #  double underscore would be name-mangled)
poison.argtypes = unpoison.argtypes = [C.c_void_p, C.c_size_t]


class Fenced:
    """`a` copied into the middle of a fresh buffer, k elements past 16-byte alignment, everything
    before (when the start is 8-byte aligned) and after the view poisoned."""

    def __init__(self, a, k):
        a = np.ascontiguousarray(a)
        per16 = 16 // a.itemsize
        self.buf = np.zeros(a.size + 4 * per16 + 16, a.dtype)
        base = (-(self.buf.ctypes.data // a.itemsize)) % per16 + per16  # aligned element, with room in front
        self.view = self.buf[base + k: base + k + a.size]
        self.view[:] = a
        lo, hi = self.view.ctypes.data, self.view.ctypes.data + a.nbytes
        b0, b1 = self.buf.ctypes.data, self.buf.ctypes.data + self.buf.nbytes
        self.regions = []
        if lo % 8 == 0 and lo > b0:
            self.regions.append((b0, lo - b0))
        if b1 > hi:
            self.regions.append((hi, b1 - hi))
        for addr, n in self.regions:
            poison(addr, n)

    def release(self):
        for addr, n in self.regions:
            unpoison(addr, n)


def fenced(ro, col, val, x, blocks, mis):
    f = [Fenced(val, mis[0]), Fenced(np.asarray(col, np.int32), mis[1]), Fenced(np.asarray(ro, np.int32), mis[2]),
         Fenced(x, 0)]
    rows, nnz = ro.size - 1, int(ro[-1])
    y = np.full(rows, np.nan, val.dtype)
    rc = call(f[0].view.ctypes.data, f[2].view.ctypes.data, f[1].view.ctypes.data, f[3].view.ctypes.data, y, rows, nnz,
              val.dtype, blocks, 1)
    for r in f:
        r.release()
    assert rc == 0
    return y


rng = np.random.default_rng(6)
shapes = [(1, 1, 1.0, 0.0, 0), (1, 50, 20, 0.0, 0), (17, 1, 1.0, 0.3, 0), (100, 64, 0.0, 1.0, 0),
          (3000, 300, 0.05, 0.9, 0), (1500, 2000, 9, 0.1, 2), (1, 40000, 30000, 0.0, 1), (9000, 128, 2, 0.5, 0),
          (7, 5, 2, 0.2, 0), (129, 33, 3, 0.1, 1), (1153, 50, 1, 0.0, 0), (2000, 3, 1, 0.3, 0)]
for rows, cols, mean_len, empty, longs in shapes:
    ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
    nnz = int(ro[-1])
    for dt in (np.float64, np.float32):
        for blocks, search in ((1, 1), (3, 1), (1 << 20, 1), (2, 0)):
            y = strict(ro, col, np.ones(nnz, dt), np.ones(cols, dt), blocks, search)
            assert np.array_equal(y, np.diff(ro).astype(dt)), (rows, cols, blocks, search)
for rows, cols, mean_len, empty, longs in [(700, 900, 7, 0.2, 1), (5, 7, 2, 0.0, 0), (1, 9, 5, 0.0, 0), (2500, 40, 1, 0.4, 0)]:
    ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
    nnz = int(ro[-1])
    for dt in (np.float64, np.float32):
        per16 = 16 // np.dtype(dt).itemsize
        for kv in range(per16):
            for kc, kr in ((0, 0), (1, 2), (2, 3), (3, 1), (2, 2)):
                y = fenced(ro, col, np.ones(nnz, dt), np.ones(cols, dt), 2, (kv, kc, kr))
                assert np.array_equal(y, np.diff(ro).astype(dt)), (rows, cols, kv, kc, kr)
print("asan check complete")
