"""The tile engine's kernel *logic* on a box without a GPU.

tests/emu/ holds a small host-side SIMT interpreter (coroutine per CUDA thread, warp collectives,
mbarrier / cp.async.bulk contracts).  It compiles merge-spmv_b200/csrc/spmv_tile.cuh UNCHANGED with
g++ -- only the inline-PTX wrappers of ptx_sm100.cuh are swapped for host stand-ins -- and runs the
search / tile / carry fix-up kernels block by block.  This is test infrastructure: the library
under test here is tests/emu/_build/libtile_emu.so, built by this file; the product
(libmergespmv.so) contains no host path and still fails loudly without a GPU.  The parity tests
proper are tests/test_gpu_parity.py (-m gpu).

What it buys: bitmap / popcount-prefix / walk / segmented-scan / fix-up bugs, staging-range
arithmetic (alignment contracts of cp.async.bulk are asserted), reads of staged data before the
mbarrier wait (the destination is poisoned until the phase completes) and barrier deadlocks are
caught by `pytest -m "not gpu"`.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, random_csr

EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "merge-spmv_b200", "csrc")
VARIANTS = {
    # name -> extra -D flags (kernel variants that exist as compile-time switches in spmv_tile.cuh)
    "shipped": [],
    "prefix_shfl": ["-DMSPMV_PREFIX_SHFL"],
    "ipt_11_15": ["-DMSPMV_TILE_IPT=(sizeof(T)==8?11:15)"],
    "ipt_7_11": ["-DMSPMV_TILE_IPT=(sizeof(T)==8?7:11)"],
    "ipt_8_12": ["-DMSPMV_TILE_IPT=(sizeof(T)==8?8:12)"],
    "v3_popcount_prefix": ["-DMSPMV_V3_XS_SCATTER=0", "-DMSPMV_V3_BALLOT_SCAN=0"],
}


def build(name, extra=()):
    out_dir = os.path.join(EMU, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libtile_emu_{name}.so")
    srcs = [os.path.join(EMU, f) for f in ("tile_emu.cpp", "simt_emu.hpp", "ptx_emu.cuh")]
    srcs += [os.path.join(CSRC, f) for f in ("spmv_tile.cuh", "spmv_tile3.cuh", "merge_common.cuh", "tma_stage.cuh",
                                             "ptx_sm100.cuh", "carry_exchange.cuh", "spmv_stream.cuh")]
    if os.path.exists(out) and all(os.path.getmtime(out) > os.path.getmtime(s) for s in srcs + [__file__]):
        return out
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fno-strict-aliasing", "-fno-gnu-unique", "-ffp-contract=off", "-fPIC", "-shared", "-w",
           "-I", EMU, "-I", CSRC, "-I", "/usr/local/cuda/include", *VARIANTS.get(name, []), *extra,
           os.path.join(EMU, "tile_emu.cpp"), "-o", out]
    subprocess.run(cmd, check=True)
    return out


class Emu:
    def __init__(self, name="shipped"):
        self.lib = C.CDLL(build(name))
        for sfx, fp in (("f64", C.c_double), ("f32", C.c_float)):
            f = getattr(self.lib, "emu_csrmv_" + sfx)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, fp, fp, C.c_int, C.c_int, C.c_void_p, C.c_int]
            g = getattr(self.lib, "emu_csrmv_stream_" + sfx)
            g.restype = C.c_int
            g.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, fp, fp, C.c_int, C.c_int, C.c_void_p]
        self.lib.emu_merge_path_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]

    def csrmv(self, ro, col, val, x, y_in=None, alpha=1.0, beta=0.0, axpby=False, misalign=(0, 0, 0),
              prefetch_ahead=0, fused=False, variant=2, stream_sms=0):
        """misalign = element offsets (values, col, row_offsets) of the array bases from 16 bytes."""
        dt = val.dtype
        rows, nnz = ro.size - 1, int(ro[-1])

        def place(a, k, dtype):  # copy `a` into a fresh buffer whose data starts k elements past 16-byte alignment
            a = np.ascontiguousarray(a, dtype=dtype)
            per16 = 16 // a.itemsize
            buf = np.empty(a.size + 2 * per16 + 8, dtype=dtype)
            base = (-(buf.ctypes.data // a.itemsize)) % per16  # first 16-byte aligned element
            view = buf[base + k: base + k + a.size]
            view[:] = a
            assert a.size == 0 or (view.ctypes.data % 16) == (k * a.itemsize) % 16
            return buf, view

        keep = []
        pv, v = place(val, misalign[0], dt)
        pc, c = place(col, misalign[1], np.int32)
        pr, r = place(ro, misalign[2], np.int32)
        keep += [pv, pc, pr]
        xx = np.ascontiguousarray(x, dtype=dt)
        y = np.full(rows, np.nan, dtype=dt) if y_in is None else np.array(y_in, dtype=dt)
        stats = np.zeros(4, np.int32)
        fn = getattr(self.lib, "emu_csrmv_" + ("f64" if dt == np.float64 else "f32"))
        mode = 1 if fused else (3 if variant == 3 else 0)
        if stream_sms:  # the alternative "stream" engine on a pretend device with that many SMs
            fn = getattr(self.lib, "emu_csrmv_stream_" + ("f64" if dt == np.float64 else "f32"))
            rc = fn(v.ctypes.data, r.ctypes.data, c.ctypes.data, xx.ctypes.data, y.ctypes.data, rows, nnz,
                    alpha, beta, int(axpby), stream_sms, stats.ctypes.data)
        else:
            rc = fn(v.ctypes.data, r.ctypes.data, c.ctypes.data, xx.ctypes.data, y.ctypes.data, rows, nnz,
                    alpha, beta, int(axpby), prefetch_ahead, stats.ctypes.data, mode)
        assert rc == 0
        self.stats = stats
        return y

    def search(self, ro, diags):
        ro = np.ascontiguousarray(ro, np.int32)
        d = np.ascontiguousarray(diags, np.int32)
        out = np.zeros((d.size, 2), np.int32)
        self.lib.emu_merge_path_search(ro.ctypes.data, ro.size - 1, int(ro[-1]), d.ctypes.data, d.size,
                                       out.ctypes.data)
        return out


@pytest.fixture(scope="module", params=list(VARIANTS))
def emu(request):
    return Emu(request.param)


@pytest.fixture(scope="module")
def emu0():
    return Emu("shipped")


def tol_for(ro, dt):
    lens = np.diff(ro).astype(np.float64)
    return np.full(lens.shape, 1e-10) if dt == np.float64 else np.maximum(1e-6, 4 * np.sqrt(lens) * 2.0 ** -24)


def assert_close(got, want, ro, dt, what=""):
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bound = tol_for(ro, dt) * np.maximum(np.abs(want.astype(np.float64)), 1e-300)
    bad = np.nonzero(~(err <= bound))[0]
    assert bad.size == 0, f"{what}: {bad.size} rows out of tolerance, first {bad[:3]}, got {got[bad[:3]]} want {want[bad[:3]]}"


SHAPES = [  # rows, cols, mean_len, empty_frac, long_rows
    (1, 1, 1.0, 0.0, 0), (1, 50, 20, 0.0, 0), (2, 1, 0.6, 0.0, 0), (17, 1, 1.0, 0.3, 0),
    (100, 64, 0.0, 1.0, 0),        # no nonzeros at all
    (3000, 300, 0.05, 0.9, 0),     # almost all rows empty: > ROWCAP row ends per tile (row offsets read from global)
    (1500, 2000, 9, 0.1, 2),       # a few long rows
    (40, 30000, 700, 0.0, 2),      # rows spanning several tiles
    (1, 40000, 30000, 0.0, 1),     # one huge row: > 256 tiles, second fix-up level
    (9000, 128, 2, 0.5, 0),        # short rows
    (1031, 1031, 31, 0.01, 0),
]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_emu_random_structures_vs_oracle(emu, orc, dt):
    rng = np.random.default_rng(2024)
    for rows, cols, mean_len, empty, longs in SHAPES:
        ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
        nnz = int(ro[-1])
        y = emu.csrmv(ro, col, np.ones(nnz, dt), np.ones(cols, dt))
        assert np.array_equal(y, np.diff(ro).astype(dt)), (rows, cols, "exact-integer inputs must be bit-exact")
        val = (0.5 + rng.random(nnz)).astype(dt)
        x = (0.5 + rng.random(cols)).astype(dt)
        want = orc.merge_csrmv(ro, col, val, x, num_threads=8)
        got = emu.csrmv(ro, col, val, x)
        assert_close(got, want, ro, dt, f"{rows}x{cols}")


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_emu_misaligned_bases(emu0, orc, dt):
    """Array bases at every element offset inside a 16-byte granule (slices of larger arrays, as the
    multi-GPU shards produce): the superset staging must stay inside the arrays and land every
    element."""
    rng = np.random.default_rng(5)
    ro, col = random_csr(rng, 700, 900, 7, 0.2, 1)
    nnz = int(ro[-1])
    val = (0.5 + rng.random(nnz)).astype(dt)
    x = (0.5 + rng.random(900)).astype(dt)
    want = orc.merge_csrmv(ro, col, val, x, num_threads=3)
    per16 = 16 // np.dtype(dt).itemsize
    for kv in range(per16):
        for kc in (0, 1, 3):
            for kr in (0, 2, 3):
                got = emu0.csrmv(ro, col, val, x, misalign=(kv, kc, kr))
                assert_close(got, want, ro, dt, f"misalign {(kv, kc, kr)}")


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_emu_axpby_and_prefetch(emu0, orc, dt):
    rng = np.random.default_rng(9)
    ro, col = random_csr(rng, 2500, 1000, 6, 0.2, 1)
    nnz = int(ro[-1])
    val = (0.5 + rng.random(nnz)).astype(dt)
    x = (0.5 + rng.random(1000)).astype(dt)
    y0 = rng.random(2500).astype(dt)
    ax = orc.merge_csrmv(ro, col, val, x, num_threads=4)
    for alpha, beta in ((1.0, 0.0), (2.5, 0.0), (1.0, 1.0), (-0.75, 0.5)):
        got = emu0.csrmv(ro, col, val, x, y_in=y0, alpha=alpha, beta=beta, axpby=True)
        want = (dt(alpha) * ax + dt(beta) * y0).astype(dt)
        err = np.abs(got - want)
        scale = np.abs(dt(alpha) * ax) + np.abs(dt(beta) * y0)
        assert np.all(err <= (1e-10 if dt == np.float64 else 3e-6) * scale), (alpha, beta)
    # the optional L2 prefetch path (MSPMV_TILE_PREFETCH) is a pure hint; its ranges must be legal
    got = emu0.csrmv(ro, col, val, x, prefetch_ahead=3)
    assert_close(got, ax, ro, dt, "prefetch_ahead")


def test_emu_known_answers(emu):
    for dt in (np.float32, np.float64):
        ro = np.array([0, 2, 2, 4, 8], np.int32)  # paper Fig. 8
        val = np.array([1, 1, 3, 3, 4, 4, 4, 4], dt)
        col = np.array([0, 2, 2, 3, 0, 1, 2, 3], np.int32)
        assert emu.csrmv(ro, col, val, np.ones(4, dt)).tolist() == [2, 0, 6, 16]


def test_emu_golden_vectors(emu0):
    g = np.load(os.path.join(GOLDEN, "merge_csrmv_ref.npz"))
    done = 0
    for c in range(int(g["num_cases"])):
        ro, col = g[f"c{c}_row_offsets"], g[f"c{c}_col"]
        if ro.size - 1 + int(ro[-1]) > 400000:  # the interpreter runs ~1M merge items per second
            continue
        for tag, dt in (("f64", np.float64), ("f32", np.float32)):
            got = emu0.csrmv(ro, col, g[f"c{c}_val_{tag}"], g[f"c{c}_x_{tag}"])
            assert_close(got, g[f"c{c}_y_{tag}_p8"], ro, dt, f"golden case {c}")
        assert np.array_equal(emu0.search(ro, g[f"c{c}_diags"]), g[f"c{c}_coords"]), f"golden coords case {c}"
        done += 1
    assert done >= 2


def test_emu_deterministic(emu0):
    rng = np.random.default_rng(11)
    ro, col = random_csr(rng, 900, 5000, 40, 0.05, 2)
    nnz = int(ro[-1])
    val = rng.random(nnz).astype(np.float32)
    x = rng.random(5000).astype(np.float32)
    a = emu0.csrmv(ro, col, val, x)
    b = emu0.csrmv(ro, col, val, x, misalign=(1, 2, 3))
    assert np.array_equal(a, b), "same decomposition, same summation order: bits must not depend on alignment"


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_emu_fused_single_launch(emu0, orc, dt):
    """The single-launch path for small matrices (mspmv_set_option("small_fused_tiles", n)): every
    block searches its own coordinates, the last block folds all carries with a chunked segmented
    scan.  Same tiles and carries as the three-launch path."""
    rng = np.random.default_rng(77)
    for rows, cols, mean_len, empty, longs in SHAPES:
        ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
        nnz = int(ro[-1])
        y = emu0.csrmv(ro, col, np.ones(nnz, dt), np.ones(cols, dt), fused=True)
        assert np.array_equal(y, np.diff(ro).astype(dt)), (rows, cols, "exact-integer inputs must be bit-exact")
        val = (0.5 + rng.random(nnz)).astype(dt)
        x = (0.5 + rng.random(cols)).astype(dt)
        got = emu0.csrmv(ro, col, val, x, fused=True)
        assert_close(got, orc.merge_csrmv(ro, col, val, x, num_threads=8), ro, dt, f"{rows}x{cols}")
    # alpha / beta through the fused path
    ro, col = random_csr(rng, 900, 400, 5, 0.2, 1)
    nnz = int(ro[-1])
    val = (0.5 + rng.random(nnz)).astype(dt)
    x = (0.5 + rng.random(400)).astype(dt)
    y0 = rng.random(900).astype(dt)
    ax = orc.merge_csrmv(ro, col, val, x, num_threads=4)
    got = emu0.csrmv(ro, col, val, x, y_in=y0, alpha=-0.75, beta=0.5, axpby=True, fused=True)
    want = (dt(-0.75) * ax + dt(0.5) * y0).astype(dt)
    assert np.all(np.abs(got - want) <= (1e-10 if dt == np.float64 else 3e-6) * (np.abs(0.75 * ax) + np.abs(0.5 * y0)))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_emu_tile_variant3_bit_identical(emu, orc, dt):
    """tile_body_v3 (thread-blocked gathers, products in registers) performs the same floating-point
    operations in the same order as tile_body: the results must be bit-identical, for every
    structure, alignment and for alpha/beta."""
    rng = np.random.default_rng(33)
    for rows, cols, mean_len, empty, longs in SHAPES:
        ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
        nnz = int(ro[-1])
        val = (0.5 + rng.random(nnz)).astype(dt)
        x = (0.5 + rng.random(cols)).astype(dt)
        v2 = emu.csrmv(ro, col, val, x)
        v3 = emu.csrmv(ro, col, val, x, variant=3)
        assert np.array_equal(v2, v3), (rows, cols)
        assert_close(v3, orc.merge_csrmv(ro, col, val, x, num_threads=8), ro, dt, f"{rows}x{cols}")
    ro, col = random_csr(rng, 700, 900, 7, 0.2, 1)
    nnz = int(ro[-1])
    val = (0.5 + rng.random(nnz)).astype(dt)
    x = (0.5 + rng.random(900)).astype(dt)
    y0 = rng.random(700).astype(dt)
    for mis in ((0, 0, 0), (1, 2, 3), (1, 1, 0), (0, 3, 1)):
        a = emu.csrmv(ro, col, val, x, y_in=y0, alpha=1.5, beta=-0.5, axpby=True, misalign=mis)
        b = emu.csrmv(ro, col, val, x, y_in=y0, alpha=1.5, beta=-0.5, axpby=True, misalign=mis, variant=3)
        assert np.array_equal(a, b), mis


@pytest.mark.parametrize("use_f32", [0, 1])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_emu_nvlink_carry_exchange(emu0, world, use_f32):
    """carry_exchange_kernel (peer-memory push + flag, then wait + fold) with `world` simulated
    ranks over several steps: slot indexing, the epoch-parity double buffer, the fold order of
    cpu_spmv.cpp:348-352 (a row spanning several ranks takes every carry; carries of rows >= rows
    are dropped; the last rank's carry is never applied)."""
    rng = np.random.default_rng(world)
    steps = 5
    rows = 40
    cuts = np.zeros(world + 1, np.int32)
    cuts[1:-1] = np.sort(rng.integers(0, rows + 1, world - 1))
    cuts[-1] = rows
    if world == 8:
        cuts[3] = cuts[4] = cuts[5]  # ranks 3 and 4 own no row: one long row spans them
    carry_rows = np.ascontiguousarray(cuts[1:]).astype(np.int32)  # rank g's carry belongs to global row cuts[g+1]
    carries = rng.integers(1, 50, (steps, world)).astype(np.float64)
    y = np.zeros((steps, rows), np.float64)
    lib = emu0.lib
    lib.emu_exchange_f64.restype = C.c_int
    lib.emu_exchange_f64.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = lib.emu_exchange_f64(world, steps, cuts.ctypes.data, carry_rows.ctypes.data, carries.ctypes.data,
                              y.ctypes.data, use_f32)
    assert rc == 0, "every rank must have advanced its epoch counter once per step"
    for st in range(steps):
        want = 100.0 * st + np.arange(rows, dtype=np.float64)
        for g in range(world - 1):
            r = int(carry_rows[g])
            if r < rows:
                want[r] += carries[st, g]
        assert np.array_equal(y[st], want), (st, cuts.tolist())


@pytest.mark.parametrize("schedule", [0, 1, 2])
def test_emu_fuzz_small_matrices(emu0, orc, schedule):
    """400 random small matrices (empty rows, single long rows, constant rows, sparse patterns),
    random base-pointer misalignment, both value types; shipped kernel, variant 3, the fused
    single-launch path and the stream engine.  Small-integer inputs make the result exact in any summation order, so the
    comparison with SpmvGold (cpu_spmv.cpp:257-277) is bit for bit."""
    # schedule: order in which the interpreter resumes the threads of a block (ascending, descending,
    # random per pass) -- a kernel whose shared-memory reads are properly ordered after the writes by
    # barriers gives the same bits under all of them
    emu0.lib.emu_set_schedule(schedule)
    rng = np.random.default_rng(20261017 + schedule)
    for it in range(400 if schedule == 0 else 200):
        dt = (np.float64, np.float32)[it & 1]
        rows = int(rng.integers(1, 400)) if rng.random() < 0.8 else int(rng.integers(400, 2500))
        cols = int(rng.integers(1, 200))
        mode = int(rng.integers(0, 5))
        if mode == 0:
            lens = rng.integers(0, min(cols, 4) + 1, rows)
        elif mode == 1:
            lens = rng.poisson(rng.uniform(0.1, 12), rows)
        elif mode == 2:
            lens = np.zeros(rows, np.int64)
            for _ in range(int(rng.integers(1, 4))):
                lens[rng.integers(rows)] = rng.integers(0, cols + 1)
        elif mode == 3:
            lens = np.full(rows, rng.integers(0, min(cols, 16) + 1))
        else:
            lens = (rng.random(rows) < rng.uniform(0.05, 0.9)) * rng.integers(1, min(cols, 30) + 1, rows)
        lens = np.minimum(lens, cols).astype(np.int64)
        ro = np.zeros(rows + 1, np.int32)
        ro[1:] = np.cumsum(lens)
        nnz = int(ro[-1])
        col = np.empty(nnz, np.int32)
        for r in np.nonzero(lens)[0]:
            col[ro[r]:ro[r + 1]] = np.sort(rng.choice(cols, lens[r], replace=False))
        val = rng.integers(1, 8, nnz).astype(dt)
        x = rng.integers(1, 8, cols).astype(dt)
        want = orc.spmv_gold(ro, col, val.astype(np.float64), x.astype(np.float64)).astype(dt)
        per16 = 16 // np.dtype(dt).itemsize
        mis = (int(rng.integers(per16)), int(rng.integers(4)), int(rng.integers(4)))
        variant = 3 if rng.random() < 0.4 else 2
        fused = variant == 2 and rng.random() < 0.3
        stream_sms = int(rng.integers(1, 5)) if rng.random() < 0.15 else 0  # now and then the stream engine instead
        got = emu0.csrmv(ro, col, val, x, misalign=mis, variant=variant, fused=fused, stream_sms=stream_sms)
        assert np.array_equal(got, want), (it, rows, cols, nnz, mode, mis, variant, fused, stream_sms, schedule)
    emu0.lib.emu_set_schedule(0)


@pytest.mark.parametrize("dt,tile", [(np.float64, 1152), (np.float32, 1664)])
def test_emu_tile_boundary_cases(emu0, orc, dt, tile):
    """Structures built around the tile size: row ends on the last / first item of a tile, two ends
    across a boundary, tiles made only of empty rows (more row ends than ROWCAP), a row spanning
    several tiles, trailing empty rows, degenerate sizes -- shipped kernel, variant 3 and the fused
    launch, aligned and misaligned bases, bit-exact against SpmvGold on small-integer inputs."""
    cases = [
        ([tile - 1] * 3, 2000), ([tile] * 3, 2000), ([tile - 2, 0] * 3, 2000), ([0] * (tile * 2), 5),
        ([0] * (tile * 2 + 1), 5), ([1] * (tile // 2 * 3), 7), ([0] * 500 + [1900] + [0] * 700, 2000),
        ([3] * 100 + [0] * tile, 10), ([tile * 3 + 5], 5000), ([2] * 5, 3), ([0], 1), (list(np.arange(60) % 9), 9),
    ]
    rng = np.random.default_rng(3)
    for lens, cols in cases:
        lens = np.asarray(lens, np.int64)
        ro = np.zeros(lens.size + 1, np.int32)
        ro[1:] = np.cumsum(lens)
        nnz = int(ro[-1])
        col = np.empty(nnz, np.int32)
        for r in np.nonzero(lens)[0]:
            col[ro[r]:ro[r + 1]] = np.sort(rng.choice(cols, lens[r], replace=False))
        val = rng.integers(1, 5, nnz).astype(dt)
        x = rng.integers(1, 5, cols).astype(dt)
        want = orc.spmv_gold(ro, col, val.astype(np.float64), x.astype(np.float64)).astype(dt)
        for variant, fused in ((2, False), (3, False), (2, True)):
            for mis in ((0, 0, 0), (1, 3, 1)):
                got = emu0.csrmv(ro, col, val, x, misalign=mis, variant=variant, fused=fused)
                assert np.array_equal(got, want), (lens.size, nnz, variant, fused, mis)


def test_emu_address_sanitizer():
    """No kernel of the tile engine touches memory outside the caller's arrays or its own temporaries:
    the interpreter build is instrumented with AddressSanitizer and run on exact-size arrays (see
    tests/emu/asan_check.py).  This is how the unguarded read of carry_rows[i-1] by the idle threads
    of the last fix-up block was found (inside the temp blob in the product, hence invisible to
    compute-sanitizer)."""
    import shutil
    import sys

    gxx = shutil.which("g++")
    libasan = subprocess.run([gxx, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan.so not available")
    lib = build("asan", extra=["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                               "-fno-omit-frame-pointer"])  # undefined behaviour (shifts, overflow, alignment) aborts too
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, os.path.join(EMU, "asan_check.py"), lib], capture_output=True, text=True,
                       env=env, timeout=900)
    assert r.returncode == 0 and "asan check complete" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])


def test_emu_thread_sanitizer():
    """Data-race check of the kernels' shared-memory protocol.  tests/emu/tsan_check.cpp builds the
    interpreter with -fsanitize=thread -DEMU_TSAN: every CUDA thread is a TSan fiber, __syncthreads /
    named barriers / warp collectives / mbarrier phases are the only release-acquire edges, atomics are
    atomics, the bytes a bulk copy lands are ordinary writes.  Any access of the shipped kernel,
    variant 3, the fused launch or the carry exchange that is not ordered by those edges is reported
    (removing the barrier after the row-end scatter, or the mbarrier wait, gives dozens of reports)."""
    import shutil

    gxx = shutil.which("g++")
    libtsan = subprocess.run([gxx, "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libtsan):
        pytest.skip("libtsan not available")
    out_dir = os.path.join(EMU, "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "tsan_check")
    subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fno-strict-aliasing", "-ffp-contract=off", "-w",
                    "-fsanitize=thread", "-DEMU_TSAN", "-I", EMU, "-I", CSRC, "-I", "/usr/local/cuda/include",
                    os.path.join(EMU, "tsan_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66"))
    if "unexpected memory mapping" in r.stderr or "FATAL: ThreadSanitizer" in r.stderr:
        pytest.skip("ThreadSanitizer cannot start in this environment: " + r.stderr[-200:])
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[:4000]
    assert r.returncode == 0 and "tsan check complete" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_emu_stream_engine(emu0, orc, dt):
    """The alternative "stream" engine (MSPMV_ENGINE=stream: persistent swaths, TMA producer warps ->
    cp.async gather warps -> reduce warps over mbarrier rings) under the interpreter, on pretend
    devices of 1 and 3 SMs, aligned (vectorised gather stage) and misaligned bases, alpha/beta."""
    rng = np.random.default_rng(88)
    for rows, cols, mean_len, empty, longs in SHAPES:
        ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
        nnz = int(ro[-1])
        for sms, mis in ((1, (0, 0, 0)), (3, (0, 0, 0)), (2, (1, 2, 3))):
            y = emu0.csrmv(ro, col, np.ones(nnz, dt), np.ones(cols, dt), stream_sms=sms, misalign=mis)
            assert np.array_equal(y, np.diff(ro).astype(dt)), (rows, cols, sms, mis, "exact")
        val = (0.5 + rng.random(nnz)).astype(dt)
        x = (0.5 + rng.random(cols)).astype(dt)
        got = emu0.csrmv(ro, col, val, x, stream_sms=3)
        assert_close(got, orc.merge_csrmv(ro, col, val, x, num_threads=8), ro, dt, f"{rows}x{cols}")
    ro, col = random_csr(rng, 900, 400, 5, 0.2, 1)
    nnz = int(ro[-1])
    val = (0.5 + rng.random(nnz)).astype(dt)
    x = (0.5 + rng.random(400)).astype(dt)
    y0 = rng.random(900).astype(dt)
    ax = orc.merge_csrmv(ro, col, val, x, num_threads=4)
    got = emu0.csrmv(ro, col, val, x, y_in=y0, alpha=-0.75, beta=0.5, axpby=True, stream_sms=2)
    want = (dt(-0.75) * ax + dt(0.5) * y0).astype(dt)
    assert np.all(np.abs(got - want) <= (1e-10 if dt == np.float64 else 3e-6) * (np.abs(0.75 * ax) + np.abs(0.5 * y0)))
