"""The pipe engine's kernel *logic* (merge-spmv_b200/csrc/spmv_pipe.cuh, the default engine) on a box
without a GPU, under the host SIMT interpreter of tests/emu/ (test infrastructure only -- see
tests/test_kernel_emu.py for what the interpreter models).  The kernel source is compiled UNCHANGED
with g++; only the inline-PTX wrappers are swapped for host stand-ins that keep the instructions'
contracts (cp.async.bulk alignment, data invisible until the mbarrier phase completes, barriers that
nobody can complete are deadlocks).

Covered here: the producer's in-kernel merge-path walk (bit-exact coordinates against the oracle),
the TMA ring (full / empty mbarriers, 2 and 3 stages), the slot walk with parked segment sums, the
block-wide segmented scan with the carry kept in registers across the tiles of a block, the
row-owner stores, and the last-block fold of the per-block carries -- for 1, few and many blocks.
The parity tests proper are tests/test_gpu_parity.py (-m gpu).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, random_csr

EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "merge-spmv_b200", "csrc")
VARIANTS = {  # <IPT fp64, IPT fp32, value-ring slots, column-ring slots, gather-ahead, consumer warps, first-in-register>
    "shipped": ["-DEMU_PIPE_CFG=9,13,2,2,0,4,0"],     # shape B of the product (short rows)
    "shipped_a": ["-DEMU_PIPE_CFG=11,13,2,1,0,4,1"],  # shape A of the product (long rows)
    "ahead_2_2": ["-DEMU_PIPE_CFG=9,13,2,2,1,4,1"],
    "ahead_v2_c1_ipt_7_11": ["-DEMU_PIPE_CFG=7,11,2,1,1,4,0"],
    "v3_c1_ipt_5_5": ["-DEMU_PIPE_CFG=5,5,3,1,0,4,1"],
    "ahead_nw2_v2_c3": ["-DEMU_PIPE_CFG=9,7,2,3,1,2,0"],
}


def build(name, extra=()):
    out_dir = os.path.join(EMU, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libpipe_emu_{name}.so")
    srcs = [os.path.join(EMU, f) for f in ("pipe_emu.cpp", "simt_emu.hpp", "ptx_emu.cuh")]
    srcs += [os.path.join(CSRC, f) for f in ("spmv_pipe.cuh", "merge_search.cuh", "carry_exchange.cuh", "merge_common.cuh", "tma_stage.cuh",
                                             "ptx_sm100.cuh")]
    if os.path.exists(out) and all(os.path.getmtime(out) > os.path.getmtime(s) for s in srcs + [__file__]):
        return out
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fno-strict-aliasing", "-fno-gnu-unique", "-ffp-contract=off", "-fPIC",
           "-shared", "-w", "-I", EMU, "-I", CSRC, "-I", "/usr/local/cuda/include", *VARIANTS.get(name, []), *extra,
           os.path.join(EMU, "pipe_emu.cpp"), "-o", out]
    subprocess.run(cmd, check=True)
    return out


class PipeEmu:
    def __init__(self, name="shipped", path=None):
        self.lib = C.CDLL(path or build(name))
        for sfx, fp in (("f64", C.c_double), ("f32", C.c_float)):
            f = getattr(self.lib, "emu_pipe_" + sfx)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, fp, fp, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.emu_merge_path_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]

    def csrmv(self, ro, col, val, x, y_in=None, alpha=1.0, beta=0.0, axpby=False, misalign=(0, 0, 0), blocks=3,
              search=True, want_coords=False):
        """misalign = element offsets (values, col, row_offsets) of the array bases from 16 bytes;
        blocks = the grid the host would launch (min'ed with the tile count like the product)."""
        dt = val.dtype
        rows, nnz = ro.size - 1, int(ro[-1])

        def place(a, k, dtype):  # exact-size view whose data starts k elements past 16-byte alignment
            a = np.ascontiguousarray(a, dtype=dtype)
            per16 = 16 // a.itemsize
            buf = np.empty(a.size + 2 * per16 + 8, dtype=dtype)
            base = (-(buf.ctypes.data // a.itemsize)) % per16
            view = buf[base + k: base + k + a.size]
            view[:] = a
            return buf, view

        pv, v = place(val, misalign[0], dt)
        pc, c = place(col, misalign[1], np.int32)
        pr, r = place(ro, misalign[2], np.int32)
        xx = np.ascontiguousarray(x, dtype=dt)
        y = np.full(rows, np.nan, dtype=dt) if y_in is None else np.array(y_in, dtype=dt)
        stats = np.zeros(4, np.int32)
        tile = self.tile(dt)
        ntiles = max((rows + nnz + tile - 1) // tile, 1)
        coords = np.full((ntiles + 1, 2), -7, np.int32)
        fn = getattr(self.lib, "emu_pipe_" + ("f64" if dt == np.float64 else "f32"))
        rc = fn(v.ctypes.data, r.ctypes.data, c.ctypes.data, xx.ctypes.data, y.ctypes.data, rows, nnz, alpha, beta,
                int(axpby), blocks, int(search), coords.ctypes.data if want_coords else None, stats.ctypes.data)
        assert rc == 0
        self.stats = stats
        if rows > 0:
            assert stats[1] == tile, "test's idea of the tile size is stale"
        self.coords = coords
        return y

    _ipt = None

    def ipt(self, dt):
        if self._ipt is None:
            stats = np.zeros(4, np.int32)
            one = np.ones(1, np.float64)
            ro = np.array([0, 1], np.int32)
            z = np.zeros(1, np.int32)
            y = np.zeros(1, np.float64)
            self.lib.emu_pipe_f64(one.ctypes.data, ro.ctypes.data, z.ctypes.data, one.ctypes.data, y.ctypes.data, 1, 1,
                                  1.0, 0.0, 0, 1, 1, None, stats.ctypes.data)
            t64 = int(stats[1])
            one32 = np.ones(1, np.float32)
            y32 = np.zeros(1, np.float32)
            self.lib.emu_pipe_f32(one32.ctypes.data, ro.ctypes.data, z.ctypes.data, one32.ctypes.data, y32.ctypes.data,
                                  1, 1, 1.0, 0.0, 0, 1, 1, None, stats.ctypes.data)
            self._ipt = {np.dtype(np.float64): t64, np.dtype(np.float32): int(stats[1])}
        return self._ipt[np.dtype(dt)]

    tile = ipt  # merge items per tile of the build under test

    def search(self, ro, diags):
        ro = np.ascontiguousarray(ro, np.int32)
        d = np.ascontiguousarray(diags, np.int32)
        out = np.zeros((d.size, 2), np.int32)
        self.lib.emu_merge_path_search(ro.ctypes.data, ro.size - 1, int(ro[-1]), d.ctypes.data, d.size, out.ctypes.data)
        return out


@pytest.fixture(scope="module", params=list(VARIANTS))
def emu(request):
    return PipeEmu(request.param)


@pytest.fixture(scope="module")
def emu0():
    return PipeEmu("shipped")


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
    (3000, 300, 0.05, 0.9, 0),     # almost all rows empty: > ROWCAP row ends per tile (row offsets read through L2)
    (1500, 2000, 9, 0.1, 2),       # a few long rows
    (40, 30000, 700, 0.0, 2),      # rows spanning several tiles
    (1, 40000, 30000, 0.0, 1),     # one huge row spanning every block
    (9000, 128, 2, 0.5, 0),        # short rows
    (1031, 1031, 31, 0.01, 0),
]


def oracle_tile_coords(orc, ro, tile):
    rows, nnz = ro.size - 1, int(ro[-1])
    ntiles = max((rows + nnz + tile - 1) // tile, 1)
    return np.array([orc.merge_path_search(min(t * tile, rows + nnz), ro) for t in range(ntiles + 1)], np.int32)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_pipe_random_structures_vs_oracle(emu, orc, dt):
    rng = np.random.default_rng(2024)
    for rows, cols, mean_len, empty, longs in SHAPES:
        ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
        nnz = int(ro[-1])
        for blocks in (1, 3, 1 << 20):  # one block walks everything / runs of several tiles / one tile per block
            y = emu.csrmv(ro, col, np.ones(nnz, dt), np.ones(cols, dt), blocks=blocks, want_coords=True)
            assert np.array_equal(y, np.diff(ro).astype(dt)), (rows, cols, blocks, "exact-integer inputs must be bit-exact")
            # the coordinates the producer warp found in-kernel == the reference's MergePathSearch
            assert np.array_equal(emu.coords, oracle_tile_coords(orc, ro, emu.tile(dt))), (rows, cols, blocks)
        val = (0.5 + rng.random(nnz)).astype(dt)
        x = (0.5 + rng.random(cols)).astype(dt)
        want = orc.merge_csrmv(ro, col, val, x, num_threads=8)
        for blocks in (2, 5):
            assert_close(emu.csrmv(ro, col, val, x, blocks=blocks), want, ro, dt, f"{rows}x{cols} blocks={blocks}")
        # coordinates from tile_search_kernel instead of the in-kernel walk: same tiles, same bits
        a = emu.csrmv(ro, col, val, x, blocks=4, search=True)
        b = emu.csrmv(ro, col, val, x, blocks=4, search=False)
        assert np.array_equal(a, b), (rows, cols)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_pipe_misaligned_bases(emu0, orc, dt):
    """Array bases at every element offset inside a 16-byte granule (slices of larger arrays, as the
    multi-GPU shards produce): the superset staging must stay inside the arrays and land every element."""
    rng = np.random.default_rng(5)
    ro, col = random_csr(rng, 700, 900, 7, 0.2, 1)
    nnz = int(ro[-1])
    val = (0.5 + rng.random(nnz)).astype(dt)
    x = (0.5 + rng.random(900)).astype(dt)
    want = orc.merge_csrmv(ro, col, val, x, num_threads=3)
    per16 = 16 // np.dtype(dt).itemsize
    ref = emu0.csrmv(ro, col, val, x)
    for kv in range(per16):
        for kc in (0, 1, 3):
            for kr in (0, 2, 3):
                got = emu0.csrmv(ro, col, val, x, misalign=(kv, kc, kr))
                assert_close(got, want, ro, dt, f"misalign {(kv, kc, kr)}")
                assert np.array_equal(got, ref), "same decomposition, same summation order: bits must not depend on alignment"


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_pipe_axpby(emu0, orc, dt):
    rng = np.random.default_rng(9)
    ro, col = random_csr(rng, 2500, 1000, 6, 0.2, 1)
    nnz = int(ro[-1])
    val = (0.5 + rng.random(nnz)).astype(dt)
    x = (0.5 + rng.random(1000)).astype(dt)
    y0 = rng.random(2500).astype(dt)
    ax = orc.merge_csrmv(ro, col, val, x, num_threads=4)
    for alpha, beta in ((1.0, 0.0), (2.5, 0.0), (1.0, 1.0), (-0.75, 0.5)):
        got = emu0.csrmv(ro, col, val, x, y_in=y0, alpha=alpha, beta=beta, axpby=True, blocks=4)
        want = (dt(alpha) * ax + dt(beta) * y0).astype(dt)
        err = np.abs(got - want)
        scale = np.abs(dt(alpha) * ax) + np.abs(dt(beta) * y0)
        assert np.all(err <= (1e-10 if dt == np.float64 else 3e-6) * scale), (alpha, beta)


def test_pipe_known_answers(emu):
    for dt in (np.float32, np.float64):
        ro = np.array([0, 2, 2, 4, 8], np.int32)  # paper Fig. 8
        val = np.array([1, 1, 3, 3, 4, 4, 4, 4], dt)
        col = np.array([0, 2, 2, 3, 0, 1, 2, 3], np.int32)
        assert emu.csrmv(ro, col, val, np.ones(4, dt)).tolist() == [2, 0, 6, 16]


def test_pipe_golden_vectors(emu0):
    g = np.load(os.path.join(GOLDEN, "merge_csrmv_ref.npz"))
    done = 0
    for c in range(int(g["num_cases"])):
        ro, col = g[f"c{c}_row_offsets"], g[f"c{c}_col"]
        if ro.size - 1 + int(ro[-1]) > 400000:  # the interpreter runs ~1M merge items per second
            continue
        for tag, dt in (("f64", np.float64), ("f32", np.float32)):
            got = emu0.csrmv(ro, col, g[f"c{c}_val_{tag}"], g[f"c{c}_x_{tag}"], blocks=7)
            assert_close(got, g[f"c{c}_y_{tag}_p8"], ro, dt, f"golden case {c}")
        assert np.array_equal(emu0.search(ro, g[f"c{c}_diags"]), g[f"c{c}_coords"]), f"golden coords case {c}"
        done += 1
    assert done >= 2


def test_pipe_many_blocks_fold(emu0, orc):
    """More blocks than one fold chunk (128), long rows spanning many blocks: the last block's chunked
    segmented scan over the per-block carries, with runs crossing chunk boundaries."""
    rng = np.random.default_rng(4)
    tile = emu0.tile(np.float64)
    lens = np.concatenate([rng.integers(0, 4, 300), [tile * 150 + 17], rng.integers(0, 4, 50), [tile * 140 + 3],
                           rng.integers(0, 30, 2000)]).astype(np.int64)
    cols = int(lens.max())
    ro = np.zeros(lens.size + 1, np.int32)
    ro[1:] = np.cumsum(lens)
    nnz = int(ro[-1])
    col = np.empty(nnz, np.int32)
    for r in np.nonzero(lens)[0]:
        col[ro[r]:ro[r + 1]] = np.sort(rng.choice(cols, lens[r], replace=False))
    val = rng.integers(1, 4, nnz).astype(np.float64)
    x = rng.integers(1, 4, cols).astype(np.float64)
    want = orc.spmv_gold(ro, col, val, x)
    got = emu0.csrmv(ro, col, val, x, blocks=1 << 20)  # one tile per block
    assert emu0.stats[3] > 256, "wanted more than two fold chunks"
    assert np.array_equal(got, want)
    got = emu0.csrmv(ro, col, val, x, blocks=131)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("schedule", [0, 1, 2])
def test_pipe_fuzz_small_matrices(emu0, orc, schedule):
    """Random small matrices (empty rows, single long rows, constant rows, sparse patterns), random
    base-pointer misalignment, both value types, random grids.  Small-integer inputs make the result
    exact in any summation order, so the comparison with SpmvGold (cpu_spmv.cpp:257-277) is bit for bit.
    schedule: order in which the interpreter resumes the threads of a block -- a kernel whose
    shared-memory reads are properly ordered after the writes gives the same bits under all of them."""
    emu0.lib.emu_set_schedule(schedule)
    rng = np.random.default_rng(20261017 + schedule)
    for it in range(300 if schedule == 0 else 150):
        dt = (np.float64, np.float32)[it & 1]
        rows = int(rng.integers(1, 400)) if rng.random() < 0.8 else int(rng.integers(400, 4000))
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
        blocks = int(rng.integers(1, 6))
        search = bool(rng.random() < 0.7)
        got = emu0.csrmv(ro, col, val, x, misalign=mis, blocks=blocks, search=search)
        assert np.array_equal(got, want), (it, rows, cols, nnz, mode, mis, blocks, search, schedule)
    emu0.lib.emu_set_schedule(0)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_pipe_tile_boundary_cases(emu0, orc, dt):
    """Structures built around the tile size: row ends on the last / first slot of a tile, two ends
    across a boundary, tiles made only of empty rows (more row ends than ROWCAP), a row spanning
    several tiles, trailing empty rows, degenerate sizes; aligned and misaligned bases; bit-exact
    against SpmvGold on small-integer inputs."""
    tile = emu0.tile(dt)
    cases = [
        ([tile - 1] * 3, 2000), ([tile] * 3, 2000), ([tile - 2, 0] * 3, 2000), ([0] * (tile * 2), 5),
        ([0] * (tile * 2 + 1), 5), ([1] * (tile // 2 * 3), 7), ([0] * 500 + [1900] + [0] * 700, 2000),
        ([3] * 100 + [0] * tile, 10), ([tile * 3 + 5], 5000), ([2] * 5, 3), ([0], 1), (list(np.arange(60) % 9), 9),
        ([tile - 1, 0, 0, 1, tile - 3, 0, 2], 2000),
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
        for blocks in (1, 2, 1 << 20):
            for mis in ((0, 0, 0), (1, 3, 1)):
                got = emu0.csrmv(ro, col, val, x, misalign=mis, blocks=blocks)
                assert np.array_equal(got, want), (lens.size, nnz, blocks, mis)


def _sanitizer_lib(flag):
    import shutil

    gxx = shutil.which("g++")
    name = subprocess.run([gxx, f"-print-file-name=lib{flag}.so"], capture_output=True, text=True).stdout.strip()
    return name if os.path.isabs(name) and os.path.exists(name) else None


def test_pipe_address_sanitizer():
    """The kernel touches no memory outside the caller's arrays or its own temporaries: the interpreter
    build is instrumented with AddressSanitizer + UBSan and run on exact-size arrays (the emulator glue
    sizes the carry arrays for exactly the launched grid)."""
    import sys

    libasan = _sanitizer_lib("asan")
    if libasan is None:
        pytest.skip("libasan.so not available")
    lib = build("asan", extra=["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"])
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, os.path.join(EMU, "pipe_asan_check.py"), lib], capture_output=True, text=True,
                       env=env, timeout=900)
    assert r.returncode == 0 and "asan check complete" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])


def test_pipe_thread_sanitizer():
    """Data-race check of the kernels' shared-memory protocol.  tests/emu/tsan_check.cpp builds the
    interpreter with -fsanitize=thread -DEMU_TSAN: every CUDA thread is a TSan fiber, __syncthreads /
    named barriers / warp collectives / mbarrier phases are the only release-acquire edges, atomics are
    atomics, the bytes a bulk copy lands are ordinary writes.  Any access of the pipe kernel (every
    thread role: producer, consumers, the last block's fold) or the carry exchange that is not ordered by those edges is reported
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


@pytest.mark.parametrize("use_f32", [0, 1])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_pipe_nvlink_carry_exchange(emu0, world, use_f32):
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
