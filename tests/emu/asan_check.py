"""Run the tile-engine kernels under the SIMT interpreter with AddressSanitizer (spawned by
tests/test_kernel_emu.py with LD_PRELOAD=libasan.so): every caller array is a fresh, exact-size
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
    f = getattr(lib, "emu_csrmv_" + sfx)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, fp, fp, C.c_int, C.c_int, C.c_void_p, C.c_int]


def strict(ro, col, val, x, mode):
    ro, col = np.array(ro, np.int32, copy=True), np.array(col, np.int32, copy=True)
    val, x = np.array(val, copy=True), np.array(x, copy=True)
    rows, nnz = ro.size - 1, int(ro[-1])
    y = np.full(rows, np.nan, val.dtype)
    stats = np.zeros(4, np.int32)
    fn = lib.emu_csrmv_f64 if val.dtype == np.float64 else lib.emu_csrmv_f32
    assert fn(val.ctypes.data, ro.ctypes.data, col.ctypes.data, x.ctypes.data, y.ctypes.data, rows, nnz, 1.0, 0.0,
              0, 0, stats.ctypes.data, mode) == 0
    return y


rng = np.random.default_rng(6)
shapes = [(1, 1, 1.0, 0.0, 0), (1, 50, 20, 0.0, 0), (17, 1, 1.0, 0.3, 0), (100, 64, 0.0, 1.0, 0),
          (3000, 300, 0.05, 0.9, 0), (1500, 2000, 9, 0.1, 2), (1, 40000, 30000, 0.0, 1), (9000, 128, 2, 0.5, 0),
          (7, 5, 2, 0.2, 0), (129, 33, 3, 0.1, 1), (1153, 50, 1, 0.0, 0), (2000, 3, 1, 0.3, 0)]
for rows, cols, mean_len, empty, longs in shapes:
    ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
    nnz = int(ro[-1])
    for dt in (np.float64, np.float32):
        for mode in (0, 1, 3):  # shipped three-launch path, fused single launch, tile variant 3
            y = strict(ro, col, np.ones(nnz, dt), np.ones(cols, dt), mode)
            assert np.array_equal(y, np.diff(ro).astype(dt)), (rows, cols, mode)
print("asan check complete")
