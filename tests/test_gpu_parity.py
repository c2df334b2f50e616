"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle, the
committed golden vectors of the reference's own code, and size-independent properties at
BASELINE.json's full sizes.  Run with `pytest -m gpu` on a B200.

Tolerances (stated once, used everywhere):
  * exact-integer inputs (values = x = 1, the reference drivers' own choice, gpu_spmv.cu:521-525):
    bit-exact, any summation order;
  * random inputs, fp64: |y - y_ref| <= 1e-10 * |y_ref|;
  * random inputs, fp32: relative error <= max(1e-6, 4*sqrt(row_len)*2^-24) against the oracle
    (summation order differs between p CPU threads and the GPU decomposition), and the
    reference's own CompareResults rule (utils.h:692-742) must PASS;
  * merge-path coordinates: bit-exact.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, random_csr

import merge_spmv_b200 as ms
from merge_spmv_b200 import generators as gen
from merge_spmv_b200 import sharded
from merge_spmv_b200.csrmv import DeviceSpmv, csrmv_config

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
# Every kernel shape the library ships, and both launch modes: "A" / "B" = the two instantiations the
# dispatch picks from by average row length (forced here, so each sees every structure), "A2" = A with
# the tile coordinates from the stand-alone search kernel (two launches) instead of the producer warp.
ENGINES = ("A", "B", "A2")
_MODE = {"A": (1, 1), "B": (2, 1), "A2": (1, 0), "auto": (-1, -1)}


@pytest.fixture(autouse=True)
def _default_engine():
    yield
    set_engine("auto")


def set_engine(name):
    cfg, search = _MODE[name]
    L = ms.lib()
    assert L.mspmv_set_option(b"pipe_config", cfg) == 0
    assert L.mspmv_set_option(b"pipe_search", search) == 0


def gpu_csrmv(ro, col, val, x, **kw):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    y = ms.csrmv(t(ro), t(col), t(val), t(x), **kw)
    torch.cuda.synchronize()
    return y.cpu().numpy()


def assert_close(got, want, ro, dtype, what=""):
    lens = np.diff(ro).astype(np.float64)
    if np.dtype(dtype) == np.float64:
        tol = np.full(lens.shape, 1e-10)
    else:
        tol = np.maximum(1e-6, 4 * np.sqrt(lens) * 2.0 ** -24)
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    bound = tol * np.maximum(np.abs(want.astype(np.float64)), 1e-300)
    bad = np.nonzero(err > bound)[0]
    assert bad.size == 0, f"{what}: {bad.size} rows out of tolerance, first row {bad[:1]}, err {err[bad[:1]]}"


# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_known_answers(engine, dt):
    set_engine(engine)
    # paper Fig. 8
    ro = np.array([0, 2, 2, 4, 8], np.int32)
    val = np.array([1, 1, 3, 3, 4, 4, 4, 4], dt)
    col = np.array([0, 2, 2, 3, 0, 1, 2, 3], np.int32)
    assert gpu_csrmv(ro, col, val, np.ones(4, dt)).tolist() == [2, 0, 6, 16]
    # cub/device/device_spmv.cuh:90-123
    m = gen.grid2d(3)
    ro, col, _ = m.numpy()
    assert gpu_csrmv(ro, col, np.ones(24, dt), np.ones(9, dt)).tolist() == [2, 3, 2, 3, 4, 3, 2, 3, 2]


@pytest.mark.parametrize("engine", ENGINES)
def test_reference_generators_exact(engine):
    # closed-form generators with x = 1: y[r] = row length
    set_engine(engine)
    for m in (gen.grid2d(300), gen.grid3d(40), gen.wheel(100000), gen.dense(2048, 513)):
        for dt in (torch.float32, torch.float64):
            ro, col, val = m.numpy()
            y = gpu_csrmv(ro, col, val.astype(np.float32 if dt == torch.float32 else np.float64),
                          np.ones(m.cols, np.float32 if dt == torch.float32 else np.float64))
            assert np.array_equal(y, np.diff(ro).astype(y.dtype)), m.name


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_random_structures_vs_oracle(orc, engine, dt):
    set_engine(engine)
    rng = np.random.default_rng(101)
    shapes = [(1, 1, 1.0, 0.0, 0), (1, 50, 20, 0.0, 0), (2, 1, 0.6, 0.0, 0), (17, 1, 1.0, 0.3, 0),
              (100, 64, 0.0, 1.0, 0),          # no nonzeros at all
              (5000, 300, 0.05, 0.9, 0),       # almost all rows empty
              (3000, 4000, 9, 0.1, 2), (257, 100000, 700, 0.0, 3), (20000, 20000, 3, 0.3, 1),
              (1, 200000, 150000, 0.0, 1),     # one huge row
              (70000, 128, 2, 0.5, 0), (4099, 4099, 31, 0.01, 0)]
    for rows, cols, mean_len, empty, longs in shapes:
        ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
        nnz = int(ro[-1])
        # exact
        y = gpu_csrmv(ro, col, np.ones(nnz, dt), np.ones(cols, dt))
        assert np.array_equal(y, np.diff(ro).astype(dt)), (rows, cols, "exact")
        # random
        val = (0.5 + rng.random(nnz)).astype(dt)
        x = (0.5 + rng.random(cols)).astype(dt)
        want = orc.merge_csrmv(ro, col, val, x, num_threads=8)
        got = gpu_csrmv(ro, col, val, x)
        assert_close(got, want, ro, dt, f"{rows}x{cols}")
        if rows >= 100:  # the reference's rule allows sqrt(ulps) <= num_rows: meaningless for tiny row counts
            assert orc.compare_results(got, want) == 0
        gold64 = orc.spmv_gold(ro, col, val.astype(np.float64), x.astype(np.float64))
        assert_close(got, gold64.astype(dt), ro, dt, f"{rows}x{cols} vs fp64 gold")


@pytest.mark.parametrize("engine", ENGINES)
def test_golden_vectors(engine):
    set_engine(engine)
    g = np.load(os.path.join(GOLDEN, "merge_csrmv_ref.npz"))
    for c in range(int(g["num_cases"])):
        ro, col = g[f"c{c}_row_offsets"], g[f"c{c}_col"]
        for tag, dt in (("f64", np.float64), ("f32", np.float32)):
            got = gpu_csrmv(ro, col, g[f"c{c}_val_{tag}"], g[f"c{c}_x_{tag}"])
            for p in (1, 8, 64):
                assert_close(got, g[f"c{c}_y_{tag}_p{p}"], ro, dt, f"golden case {c} p={p}")
        rod = torch.from_numpy(ro).to(DEV)
        coords = ms.merge_path_search(rod, torch.from_numpy(g[f"c{c}_diags"])).cpu().numpy()
        assert np.array_equal(coords, g[f"c{c}_coords"]), f"golden coords case {c}"


def test_coordinates_bit_exact(orc):
    rng = np.random.default_rng(7)
    for rows, cols, mean_len, empty, longs in [(1000, 500, 5, 0.3, 2), (200000, 1000, 11, 0.05, 3),
                                               (50, 100000, 5000, 0.2, 0)]:
        ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
        nnz = int(ro[-1])
        diags = np.unique(np.concatenate([rng.integers(0, rows + nnz + 100, 4000), [0, rows + nnz]]))
        got = ms.merge_path_search(torch.from_numpy(ro).to(DEV), torch.from_numpy(diags.astype(np.int32)))
        want = np.array([orc.merge_path_search(int(d), ro) for d in diags], np.int32)
        assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("engine", ("A", "B"))
@pytest.mark.parametrize("dt,vb", [(torch.float32, 4), (torch.float64, 8)])
def test_in_kernel_swath_coordinates_bit_exact(orc, engine, dt, vb):
    """The tile coordinates the producer warps derive in-kernel (32-ary warp search for the first tile
    of a block, windowed bounded search for the rest) equal the reference's MergePathSearch on the
    tile diagonals (cpu_spmv.cpp:223-245), and equal the debug export of the stand-alone search kernel."""
    set_engine(engine)
    L = ms.lib()
    for name, scale in (("powerlaw_2m", 1 / 16), ("banded_10m", 1 / 64), ("uniform_1m_64", 1 / 32)):
        m = gen.make_config(name, scale=scale, dtype=dt).to(DEV)
        x = torch.ones(m.cols, dtype=dt, device=DEV)
        y = torch.empty(m.rows, dtype=dt, device=DEV)
        err, nbytes = DeviceSpmv.CsrMV(None, 0, None, None, None, None, None, m.rows, m.cols, m.nnz, dtype=dt)
        assert err == 0
        temp = torch.zeros(nbytes, dtype=torch.uint8, device=DEV)
        assert temp.data_ptr() % 256 == 0
        assert L.mspmv_set_option(b"pipe_export_coords", 1) == 0
        try:
            err, _ = DeviceSpmv.CsrMV(temp, nbytes, m.val, m.row_offsets, m.col, x, y, m.rows, m.cols, m.nnz)
        finally:
            L.mspmv_set_option(b"pipe_export_coords", 0)
        assert err == 0
        torch.cuda.synchronize()
        cfg = csrmv_config(vb, m.rows, m.nnz)
        t, tile = cfg["tiles"], cfg["tile_items"]
        in_kernel = temp[: (t + 1) * 8].view(torch.int32).view(-1, 2).cpu().numpy()
        ro = m.row_offsets.cpu().numpy()
        diags = np.minimum(np.arange(t + 1, dtype=np.int64) * tile, m.rows + m.nnz)
        want = np.array([orc.merge_path_search(int(d), ro) for d in diags], np.int32)
        assert np.array_equal(in_kernel, want), name
        assert np.array_equal(ms.swath_coords(m.row_offsets, vb).cpu().numpy(), want), name


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name,scale", [("cpu_uniform_16k", 1.0), ("uniform_1m_64", 1 / 16),
                                        ("powerlaw_2m", 1 / 32), ("banded_10m", 1 / 16)])
def test_baseline_configs_vs_oracle(orc, engine, name, scale):
    set_engine(engine)
    for values in ("ones", "random"):
        m = gen.make_config(name, values=values, scale=scale)
        dt = m.val.numpy().dtype
        x = gen.vector(m.cols, m.val.dtype, "ones" if values == "ones" else "random").numpy()
        ro, col, val = m.numpy()
        want = orc.merge_csrmv(ro, col, val, x, num_threads=16)
        got = gpu_csrmv(ro, col, val, x)
        if values == "ones":
            assert np.array_equal(got, want)
        else:
            assert_close(got, want, ro, dt, name)
            assert orc.compare_results(got, want) == 0


def torch_reference_fp64(m, x):
    """fp64 segment-sum reference on the GPU (torch), for sizes the CPU oracle is too slow for."""
    prod = m.val.double() * x.double()[m.col.long()]
    csum = torch.zeros(m.nnz + 1, dtype=torch.float64, device=prod.device)
    torch.cumsum(prod, 0, out=csum[1:])
    ro = m.row_offsets.long()
    return csum[ro[1:]] - csum[ro[:-1]]


@pytest.mark.parametrize("name", ["uniform_1m_64", "powerlaw_2m", "banded_10m"])
def test_full_size_properties(name):
    """BASELINE.json full sizes: exact row-length identity, linearity, determinism, and agreement
    with an fp64 torch reference."""
    set_engine("auto")
    m = gen.make_config(name, values="ones", device=DEV)
    dt = m.val.dtype
    x1 = torch.ones(m.cols, dtype=dt, device=DEV)
    y = torch.full((m.rows,), float("nan"), dtype=dt, device=DEV)  # poison: every row must be written
    ms.csrmv(m.row_offsets, m.col, m.val, x1, y)
    lens = torch.diff(m.row_offsets).to(dt)
    assert torch.equal(y, lens), f"{name}: y != row lengths"
    # random values: against fp64 torch reference
    col, val = gen.fill_nonzeros(m.row_offsets.cpu(), m.cols, 0, m.nnz,
                                 kind="banded" if name.startswith("banded") else "stratified",
                                 dtype=dt, values="random", device=DEV, out_col=m.col)
    m.val = val
    x = gen.vector(m.cols, dt, "random", device=DEV)
    y1 = ms.csrmv(m.row_offsets, m.col, m.val, x).clone()
    ref = torch_reference_fp64(m, x)
    rel = (y1.double() - ref).abs() / ref.abs().clamp_min(1e-300)
    if dt == torch.float64:
        # the cumsum reference itself carries ~nnz*eps absolute error; use a local bound
        assert float(rel.max()) < 1e-6
    else:
        tol = torch.clamp(4 * torch.sqrt(lens.double()) * 2.0 ** -24, min=1e-6)
        assert bool((rel <= tol).all()), f"max rel {float(rel.max()):.3e}"
    # determinism: bitwise identical on repeat
    y2 = ms.csrmv(m.row_offsets, m.col, m.val, x)
    assert torch.equal(y1, y2)
    # linearity in x with an exactly representable scale
    y4 = ms.csrmv(m.row_offsets, m.col, m.val, x * 4)
    assert torch.equal(y4, y1 * 4)


def test_config5_full_size_exact():
    """BASELINE config 5 at full size on one GPU (fp32, 20M x 20M power-law, ~1e9 nonzeros: rows + nnz = 1.02e9,
    i.e. the int32 merge-path arithmetic near the top of its range): values = x = 1, NaN-poisoned y -- every row
    must come out as its exact length; twice, bit-identically; and the in-kernel coordinates of the last tiles
    must reach (rows, nnz) exactly."""
    set_engine("auto")
    m = gen.make_config("powerlaw_20m", values="ones", device=DEV)
    assert m.rows + m.nnz > 1_000_000_000 and m.val.dtype == torch.float32
    x1 = torch.ones(m.cols, dtype=torch.float32, device=DEV)
    y = torch.full((m.rows,), float("nan"), dtype=torch.float32, device=DEV)
    ms.csrmv(m.row_offsets, m.col, m.val, x1, y)
    lens = torch.diff(m.row_offsets).to(torch.float32)
    assert torch.equal(y, lens), "config 5: y != row lengths"
    y2 = torch.full_like(y, float("nan"))
    ms.csrmv(m.row_offsets, m.col, m.val, x1, y2)
    assert torch.equal(y, y2)
    coords = ms.swath_coords(m.row_offsets, 4)
    assert coords[-1].tolist() == [m.rows, m.nnz] and coords[0].tolist() == [0, 0]
    assert bool((coords[1:] >= coords[:-1]).all())


def test_temp_storage_protocol():
    m = gen.make_config("uniform_1m_64", scale=1 / 64).to(DEV)
    x = torch.ones(m.cols, dtype=torch.float64, device=DEV)
    y = torch.full((m.rows,), float("nan"), dtype=torch.float64, device=DEV)
    err, nbytes = DeviceSpmv.CsrMV(None, 0, None, None, None, None, None, m.rows, m.cols, m.nnz,
                                   dtype=torch.float64)
    assert err == 0 and nbytes > 0
    temp = torch.empty(nbytes + 64, dtype=torch.uint8, device=DEV)
    # too small -> cudaErrorInvalidValue, nothing written (util_device.cuh:90-93)
    err, _ = DeviceSpmv.CsrMV(temp, nbytes - 1, m.val, m.row_offsets, m.col, x, y, m.rows, m.cols, m.nnz)
    assert err == 1
    assert bool(torch.isnan(y).all())
    # unaligned blob of uninitialised garbage is fine; blob is reusable
    temp.fill_(0xA5)
    blob = temp[3:]
    for _ in range(2):
        y.fill_(float("nan"))
        err, _ = DeviceSpmv.CsrMV(blob, nbytes, m.val, m.row_offsets, m.col, x, y, m.rows, m.cols, m.nnz)
        assert err == 0
        assert torch.equal(y, torch.full_like(y, 64.0))
    # debug_synchronous path runs and syncs (dispatch_spmv_orig.cuh:579-590)
    err, _ = DeviceSpmv.CsrMV(blob, nbytes, m.val, m.row_offsets, m.col, x, y, m.rows, m.cols, m.nnz,
                              None, True)
    assert err == 0


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_alpha_beta(orc, engine, dt):
    set_engine(engine)
    rng = np.random.default_rng(5)
    ro, col = random_csr(rng, 30000, 5000, 6, 0.2, 2)
    nnz = int(ro[-1])
    val, x, yin = (0.5 + rng.random(nnz)).astype(dt), (0.5 + rng.random(5000)).astype(dt), rng.random(30000).astype(dt)
    for alpha, beta in ((1.0, 0.0), (2.5, 0.0), (1.0, 1.0), (-0.75, 0.5)):
        want = orc.spmv_gold(ro, col, val, x, yin, alpha=alpha, beta=beta)
        t = lambda a: torch.from_numpy(a).to(DEV)
        y = t(yin.copy())
        ms.csrmv(t(ro), t(col), t(val), t(x), y, alpha=alpha, beta=beta)
        got = y.cpu().numpy()
        lens = np.diff(ro)
        scale = np.abs(alpha) * orc.spmv_gold(ro, col, np.abs(val), np.abs(x)) + np.abs(beta * yin)
        tol = (1e-10 if dt == np.float64 else 4e-6) * np.maximum(scale, 1e-30)
        assert np.all(np.abs(got - want) <= tol), (alpha, beta)


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_misaligned_base_pointers(engine, dt):
    """Sub-array views whose bases are not 16-byte aligned (shards, user slices) take the same
    TMA path with ragged edges patched."""
    set_engine(engine)
    m = gen.make_config("powerlaw_2m", scale=1 / 64, dtype=dt, values="random")
    x = gen.vector(m.cols, dt, "random").to(DEV)
    base = ms.csrmv(*(t.to(DEV) for t in (m.row_offsets, m.col, m.val)), x).clone()
    for ov, oc, orow in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 3, 2), (1, 2, 3)):
        def shifted(t, off):
            buf = torch.empty(t.numel() + 8, dtype=t.dtype, device=DEV)
            view = buf[off:off + t.numel()]
            view.copy_(t)
            return view
        val, col, ro = shifted(m.val, ov), shifted(m.col, oc), shifted(m.row_offsets, orow)
        y = ms.csrmv(ro, col, val, x)
        assert torch.equal(y, base), (ov, oc, orow)


def test_non_default_stream_and_async():
    m = gen.make_config("uniform_1m_64", scale=1 / 32).to(DEV)
    x = torch.ones(m.cols, dtype=torch.float64, device=DEV)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        y = ms.csrmv(m.row_offsets, m.col, m.val, x)
    s.synchronize()
    assert torch.equal(y, torch.full_like(y, 64.0))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_session_host_buffers(orc, dt):
    m = gen.make_config("uniform_1m_64", scale=1 / 64, values="random", dtype=torch.float64)
    ro, col, val = m.numpy()
    val = val.astype(dt)
    sess = ms.SpmvSession(ro, col, val, m.cols)
    n = 5
    xs = torch.empty((n, m.cols), dtype=torch.from_numpy(val).dtype).pin_memory()
    ys = torch.empty((n, m.rows), dtype=xs.dtype).pin_memory()
    for i in range(n):
        xs[i] = gen.vector(m.cols, xs.dtype, "random", seed=100 + i)
    y0 = np.empty(m.rows, dt)
    sess.apply(xs[0].numpy(), y0)
    want0 = orc.merge_csrmv(ro, col, val, xs[0].numpy(), 8)
    assert_close(y0, want0, ro, dt, "session.apply")
    sess.apply_many(n, xs, ys)
    assert np.array_equal(ys[0].numpy(), y0)
    for i in range(n):
        assert_close(ys[i].numpy(), orc.merge_csrmv(ro, col, val, xs[i].numpy(), 8), ro, dt, f"apply_many {i}")
    sess.close()


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_sharded_single_process_emulation(orc, world, dt):
    """All `world` shards executed one after another on one GPU; the carry exchange is emulated
    by concatenating each shard's y_local[-1].  Result must equal the unsharded CsrMV and the
    oracle run with `world` threads (same decomposition => fp64 bit-exact structure)."""
    m = gen.make_config("powerlaw_2m", scale=1 / 64, dtype=dt, values="random")
    ro, col, val = m.numpy()
    x = gen.vector(m.cols, dt, "random").to(DEV)
    md = m.to(DEV)
    full = ms.csrmv(md.row_offsets, md.col, md.val, x).cpu().numpy()
    shards, ys = [], []
    for g in range(world):
        sh = sharded.make_shard(ro, m.cols, g, world, lambda k0, k1: (md.col[k0:k1], md.val[k0:k1]), DEV)
        y_local = ms.csrmv(sh.row_offsets, sh.col, sh.val, x, num_cols=m.cols).clone()
        shards.append(sh)
        ys.append(y_local)
    carries = torch.stack([y[-1] for y in ys])
    out = np.empty(m.rows, full.dtype)
    for sh, y_local in zip(shards, ys):
        sharded.apply_carries(y_local, sh, carries)
        out[sh.x0:sh.x1] = y_local[:sh.owned_rows].cpu().numpy()
    assert_close(out, full, ro, full.dtype, "sharded vs unsharded")
    assert_close(out, orc.merge_csrmv(ro, col, val, x.cpu().numpy(), world), ro, full.dtype, "sharded vs oracle")
    # exact-integer inputs: bit-exact
    ones = torch.ones_like(md.val)
    x1 = torch.ones_like(x)
    ys = [ms.csrmv(sh.row_offsets, sh.col, ones[sh.y0:sh.y1], x1, num_cols=m.cols).clone() for sh in shards]
    carries = torch.stack([y[-1] for y in ys])
    for sh, y_local in zip(shards, ys):
        sharded.apply_carries(y_local, sh, carries)
        assert np.array_equal(y_local[:sh.owned_rows].cpu().numpy(), np.diff(ro)[sh.x0:sh.x1].astype(full.dtype))


def test_launch_counter_and_config():
    L = ms.lib()
    m = gen.make_config("uniform_1m_64", scale=1 / 64).to(DEV)
    x = torch.ones(m.cols, dtype=torch.float64, device=DEV)
    before = L.mspmv_launch_count()
    ms.csrmv(m.row_offsets, m.col, m.val, x)
    cfg = csrmv_config(8, m.rows, m.nnz)
    assert L.mspmv_launch_count() - before == cfg["kernels_per_call"] == 1  # one launch per CsrMV
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert cfg["threads"] % 32 == 0 and cfg["blocks"] <= min(cfg["tiles"], 32 * sms)
    set_engine("A2")
    before = L.mspmv_launch_count()
    ms.csrmv(m.row_offsets, m.col, m.val, x)
    assert L.mspmv_launch_count() - before == csrmv_config(8, m.rows, m.nnz)["kernels_per_call"] == 2


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("layout", ["one_device_3_shards", "all_devices"])
def test_mg_session_c_abi(orc, dt, layout):
    """mspmv_mg_session_* (one process, several shards, NVLink carry-exchange kernel, no NCCL): on a single
    GPU the shards share the device (same kernels, same exchange protocol through same-device pointers); with
    more GPUs every device gets one shard.  Against the oracle with p = shards threads -- the same
    decomposition -- and bit-exact on all-ones inputs; apply_many == apply bit for bit."""
    ndev = torch.cuda.device_count()
    devices = [0, 0, 0] if layout == "one_device_3_shards" else list(range(ndev))
    if layout == "all_devices" and ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    m = gen.make_config("powerlaw_2m", scale=1 / 64, values="random", dtype=torch.float64)
    ro, col, val = m.numpy()
    val = val.astype(dt)
    tdt = torch.from_numpy(val).dtype
    sess = ms.MultiGpuSpmvSession(ro, col, val, m.cols, devices)
    cuts = [sess.shard(g) for g in range(len(devices))]
    assert cuts[0]["x0"] == 0 and cuts[-1]["x1"] == m.rows and cuts[-1]["y1"] == m.nnz
    assert [c["device"] for c in cuts] == devices
    n = 5
    xs = torch.empty((n, m.cols), dtype=tdt).pin_memory()
    ys = torch.full((n, m.rows), float("nan"), dtype=tdt).pin_memory()
    for i in range(n):
        xs[i] = gen.vector(m.cols, tdt, "random", seed=300 + i)
    y0 = torch.full((m.rows,), float("nan"), dtype=tdt).pin_memory()
    sess.apply(xs[0], y0)
    assert_close(y0.numpy(), orc.merge_csrmv(ro, col, val, xs[0].numpy(), len(devices)), ro, dt, "mg apply")
    sess.apply_many(n, xs, ys)
    assert np.array_equal(ys[0].numpy(), y0.numpy())
    for i in range(n):
        assert_close(ys[i].numpy(), orc.merge_csrmv(ro, col, val, xs[i].numpy(), len(devices)), ro, dt, f"mg apply_many {i}")
    assert sess.time_device(3) > 0
    sess.close()
    ones = ms.MultiGpuSpmvSession(ro, col, np.ones_like(val), m.cols, devices)
    y1 = torch.full((m.rows,), float("nan"), dtype=tdt).pin_memory()
    ones.apply(torch.ones(m.cols, dtype=tdt).pin_memory(), y1)
    assert np.array_equal(y1.numpy(), np.diff(ro).astype(dt))
    ones.close()
