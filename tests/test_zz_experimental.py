"""GPU parity of the opt-in variants that have NOT been timed on a B200 yet (they were written in a
session without GPU minutes; their kernel logic is covered on CPU by tests/test_kernel_emu.py).
They are off by default in the library, so these tests are off by default too:

    MSPMV_TEST_EXPERIMENTAL=1 python -m pytest tests/test_zz_experimental.py -m gpu -q

The file sorts last so that `pytest -x` reaches it only after the shipped path has passed.
Current subjects:
  * small_fused_tiles  -- the single-launch path for small matrices (spmv_tile_fused_kernel);
  * tile_variant=3     -- thread-blocked gathers with the products in registers (spmv_tile3_kernel):
                          same floating-point operations in the same order, so bit-identical to the
                          shipped kernel;
  * gpu_spmv --cub     -- the toolkit's cub::DeviceSpmv as a second same-box comparator.
"""
import os

import numpy as np
import pytest
import torch

from conftest import random_csr

import merge_spmv_b200 as ms
from merge_spmv_b200 import generators as gen
from merge_spmv_b200.csrmv import csrmv_config

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MSPMV_TEST_EXPERIMENTAL") != "1",
                                 reason="opt-in variants: set MSPMV_TEST_EXPERIMENTAL=1")]

DEV = "cuda:0"


@pytest.fixture
def fused():
    assert ms.lib().mspmv_set_option(b"small_fused_tiles", 1 << 20) == 0
    yield
    assert ms.lib().mspmv_set_option(b"small_fused_tiles", -1) == 0


def gpu_csrmv(ro, col, val, x, **kw):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    y = ms.csrmv(t(ro), t(col), t(val), t(x), **kw)
    torch.cuda.synchronize()
    return y.cpu().numpy()


def tol_for(ro, dt):
    lens = np.diff(ro).astype(np.float64)
    return np.full(lens.shape, 1e-10) if np.dtype(dt) == np.float64 else np.maximum(1e-6, 4 * np.sqrt(lens) * 2.0 ** -24)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_fused_vs_oracle_and_three_launch(orc, fused, dt):
    rng = np.random.default_rng(303)
    shapes = [(1, 1, 1.0, 0.0, 0), (1, 50, 20, 0.0, 0), (17, 1, 1.0, 0.3, 0), (100, 64, 0.0, 1.0, 0),
              (5000, 300, 0.05, 0.9, 0), (3000, 4000, 9, 0.1, 2), (257, 100000, 700, 0.0, 3),
              (20000, 20000, 3, 0.3, 1), (1, 200000, 150000, 0.0, 1), (70000, 128, 2, 0.5, 0),
              (16384, 16384, 32, 0.0, 0)]  # the last one is BASELINE config 1
    for rows, cols, mean_len, empty, longs in shapes:
        ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
        nnz = int(ro[-1])
        assert csrmv_config(np.dtype(dt).itemsize, rows, nnz)["kernels_per_call"] == 1
        y = gpu_csrmv(ro, col, np.ones(nnz, dt), np.ones(cols, dt))
        assert np.array_equal(y, np.diff(ro).astype(dt)), (rows, cols, "exact")
        val = (0.5 + rng.random(nnz)).astype(dt)
        x = (0.5 + rng.random(cols)).astype(dt)
        want = orc.merge_csrmv(ro, col, val, x, num_threads=8)
        got = gpu_csrmv(ro, col, val, x)
        err = np.abs(got.astype(np.float64) - want.astype(np.float64))
        assert np.all(err <= tol_for(ro, dt) * np.maximum(np.abs(want), 1e-300)), (rows, cols)
        # twice: same bits (deterministic fold)
        assert np.array_equal(got, gpu_csrmv(ro, col, val, x))
        y0 = (0.5 + rng.random(rows)).astype(dt)
        got_ab = gpu_csrmv(ro, col, val, x, y=torch.from_numpy(y0).to(DEV), alpha=2.0, beta=-1.0)
        ref = 2.0 * want.astype(np.float64) - y0.astype(np.float64)
        scale = 2.0 * np.abs(want.astype(np.float64)) + np.abs(y0.astype(np.float64))
        assert np.all(np.abs(got_ab - ref) <= 4 * tol_for(ro, dt) * scale), (rows, cols, "alpha/beta")


def test_fused_in_cuda_graph(fused):
    """The ticket memset is a graph node: capture once, replay many, same result every time."""
    rng = np.random.default_rng(5)
    ro, col = random_csr(rng, 16384, 16384, 32, 0.0, 0)
    nnz = int(ro[-1])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    rod, cold, vald, xd = t(ro), t(col), t((0.5 + rng.random(nnz))), t(0.5 + rng.random(16384))
    y = torch.empty(16384, dtype=torch.float64, device=DEV)
    ms.csrmv(rod, cold, vald, xd, y)  # warm (temp blob, attributes)
    torch.cuda.synchronize()
    first = y.clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ms.csrmv(rod, cold, vald, xd, y)
    for _ in range(5):
        y.fill_(float("nan"))
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(y, first)


@pytest.fixture
def variant3():
    yield lambda on: ms.lib().mspmv_set_option(b"tile_variant", 3 if on else 2)
    assert ms.lib().mspmv_set_option(b"tile_variant", -1) == 0


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_variant3_bit_identical_random_structures(orc, variant3, dt):
    rng = np.random.default_rng(404)
    shapes = [(1, 1, 1.0, 0.0, 0), (1, 50, 20, 0.0, 0), (17, 1, 1.0, 0.3, 0), (100, 64, 0.0, 1.0, 0),
              (5000, 300, 0.05, 0.9, 0), (3000, 4000, 9, 0.1, 2), (257, 100000, 700, 0.0, 3),
              (20000, 20000, 3, 0.3, 1), (1, 200000, 150000, 0.0, 1), (70000, 128, 2, 0.5, 0),
              (4099, 4099, 31, 0.01, 0)]
    for rows, cols, mean_len, empty, longs in shapes:
        ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
        nnz = int(ro[-1])
        val = (0.5 + rng.random(nnz)).astype(dt)
        x = (0.5 + rng.random(cols)).astype(dt)
        variant3(False)
        v2 = gpu_csrmv(ro, col, val, x)
        variant3(True)
        v3 = gpu_csrmv(ro, col, val, x)
        assert np.array_equal(v2, v3), (rows, cols)
        want = orc.merge_csrmv(ro, col, val, x, num_threads=8)
        err = np.abs(v3.astype(np.float64) - want.astype(np.float64))
        assert np.all(err <= tol_for(ro, dt) * np.maximum(np.abs(want), 1e-300)), (rows, cols)
        y = gpu_csrmv(ro, col, np.ones(nnz, dt), np.ones(cols, dt))
        assert np.array_equal(y, np.diff(ro).astype(dt)), (rows, cols, "exact")


@pytest.mark.parametrize("name", ["uniform_1m_64", "powerlaw_2m", "banded_10m"])
def test_variant3_bit_identical_full_size(variant3, name):
    """BASELINE.json configs 2-4 at full size: variant 3 against the shipped kernel, bit for bit."""
    kind, dt, p = gen.CONFIGS[name]
    m = gen.make_config(name, values="random", dtype=dt, device=DEV)
    x = gen.vector(m.cols, dt, "random", device=DEV)
    variant3(False)
    y2 = ms.csrmv(m.row_offsets, m.col, m.val, x).clone()
    variant3(True)
    y3 = ms.csrmv(m.row_offsets, m.col, m.val, x)
    torch.cuda.synchronize()
    assert torch.equal(y2, y3)


def test_driver_toolkit_cub_comparator():
    """gpu_spmv --cub: the CUDA toolkit's own cub::DeviceSpmv::CsrMV (the maintained descendant of the
    reference's kernels) must PASS the driver's self-check next to the merge CsrMV."""
    import subprocess

    from conftest import ROOT
    exe = os.path.join(ROOT, "merge-spmv_b200", "gpu_spmv")
    for flags in (["--uniform=64", "--rows=65536", "--values=random", "--randx"], ["--banded=3", "--rows=200000"],
                  ["--powerlaw=20000", "--rows=50000", "--nnz=3000000", "--fp32"]):
        r = subprocess.run([exe, "--i=20", "--cub", "--cusparse"] + flags, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert "CUDA toolkit cub::DeviceSpmv CsrMV" in r.stdout and "FAIL" not in r.stdout and r.stdout.count("PASS") == 3, r.stdout


def test_driver_single_process_multi_gpu():
    """gpu_spmv --gpus=N: one process drives N devices, merge-path shards, carries exchanged by
    mspmv_exchange_carries_* through peer memory (no NCCL); the gathered y must PASS the self-check."""
    import subprocess

    from conftest import ROOT
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    exe = os.path.join(ROOT, "merge-spmv_b200", "gpu_spmv")
    for flags in (["--uniform=64", "--rows=262144", "--values=random", "--randx"],
                  ["--powerlaw=200000", "--rows=100000", "--nnz=5000000", "--fp32"], ["--banded=3", "--rows=500000"]):
        r = subprocess.run([exe, "--i=50", f"--gpus={min(n, 8)}"] + flags, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert "NVLink carry exchange" in r.stdout and "FAIL" not in r.stdout and r.stdout.count("PASS") == 2, r.stdout
