"""Synthetic inputs: the reference's generators pinned to golden CSR from its own code, and the
builder-defined uniform / power-law / banded families.  CPU only."""
import os

import numpy as np
import torch

from conftest import GOLDEN

from merge_spmv_b200 import generators as gen


def test_reference_generators_match_golden():
    g = np.load(os.path.join(GOLDEN, "generators_ref.npz"))
    for kind, a, b in (("grid2d", 3, 0), ("grid2d", 6, 0), ("grid3d", 3, 0), ("grid3d", 4, 0),
                       ("wheel", 7, 0), ("wheel", 1, 0), ("dense", 4, 3), ("dense", 5, 8)):
        m = {"grid2d": lambda: gen.grid2d(a), "grid3d": lambda: gen.grid3d(a), "wheel": lambda: gen.wheel(a),
             "dense": lambda: gen.dense(a, b)}[kind]()
        key = f"{kind}_{a}_{b}"
        ro, col, val = m.numpy()
        assert [m.rows, m.cols, m.nnz] == g[f"{key}_dims"].tolist(), key
        assert np.array_equal(ro, g[f"{key}_row_offsets"]), key
        assert np.array_equal(col, g[f"{key}_col"]), key
        assert np.array_equal(val, g[f"{key}_val"]), key


def test_lattice_is_the_docstring_matrix():
    m = gen.grid2d(3)  # cub/device/device_spmv.cuh:90-123
    assert m.row_offsets.tolist() == [0, 2, 5, 7, 10, 14, 17, 19, 22, 24]
    assert m.col.tolist() == [1, 3, 0, 2, 4, 1, 5, 0, 4, 6, 1, 3, 5, 7, 2, 4, 8, 3, 7, 4, 6, 8, 5, 7]


def _check_rows_sorted_distinct(m):
    ro, col, _ = m.numpy()
    inner = np.ones(col.size, bool)
    inner[ro[:-1][np.diff(ro) > 0]] = False  # first entry of each non-empty row
    assert np.all(np.diff(col.astype(np.int64))[inner[1:]] > 0)
    assert col.min() >= 0 and col.max() < m.cols


def test_uniform_config():
    m = gen.make_config("uniform_1m_64", scale=1 / 256)
    assert m.rows == 4096 and m.nnz == 4096 * 64
    assert torch.all(torch.diff(m.row_offsets) == 64)
    _check_rows_sorted_distinct(m)
    # marginally uniform columns: chi-square-ish sanity over 16 bins
    hist = torch.histc(m.col.double(), bins=16, min=0, max=m.cols)
    assert (hist.max() - hist.min()) / hist.mean() < 0.05


def test_powerlaw_config():
    m = gen.make_config("powerlaw_2m", scale=1 / 100)
    lens = torch.diff(m.row_offsets)
    assert m.rows == 20000 and abs(m.nnz - 2_000_000) / 2e6 < 0.01
    assert int(lens.max()) == 10000 and int(lens.min()) >= 1
    _check_rows_sorted_distinct(m)


def test_banded_config():
    m = gen.make_config("banded_10m", scale=1 / 1000)
    assert m.nnz == 7 * m.rows - 12
    ro, col, _ = m.numpy()
    r = np.repeat(np.arange(m.rows), np.diff(ro))
    assert np.all(np.abs(col - r) <= 3)
    _check_rows_sorted_distinct(m)


def test_range_generation_is_consistent():
    # any nonzero sub-range reproduces the slice of the full matrix (ranks generate own shards)
    m = gen.make_config("powerlaw_2m", scale=1 / 400, values="random")
    ro = m.row_offsets
    for k0, k1 in ((0, 17), (1000, 5000), (m.nnz - 33, m.nnz)):
        col, val = gen.fill_nonzeros(ro, m.cols, k0, k1, kind="stratified", dtype=torch.float32,
                                     values="random", seed=0x5EED0003, chunk=1000)
        assert torch.equal(col, m.col[k0:k1]) and torch.equal(val, m.val[k0:k1])
    assert float(m.val.min()) >= 0.5 and float(m.val.max()) < 1.5


def test_algorithmic_bytes_config2():
    # 826.3 MB for fp64 1M x 1M, 64 nnz/row
    rows = cols = 1 << 20
    nnz = rows * 64
    assert nnz * 12 + (rows + 1) * 4 + rows * 8 + cols * 8 == 826_277_892
