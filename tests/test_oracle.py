"""Pins the CPU oracle (oracle/merge_oracle.c) -- CPU only.

Three anchors: the known answers the reference carries, the committed golden vectors produced
by the reference's own compiled code (tests/golden/make_golden.py), and -- where oracle/_ref is
present -- the reference itself, live, on fresh random structures."""
import numpy as np
import pytest

from conftest import random_csr


# ---- known answers held by the reference ---------------------------------------------------------
def test_paper_fig8_coordinates_and_y(orc):
    # merge-based-spmv-sc16-preprint.pdf Fig. 8 / merge_spmv.png
    ro = np.array([0, 2, 2, 4, 8], np.int32)
    val = np.array([1, 1, 3, 3, 4, 4, 4, 4], np.float64)
    col = np.array([0, 2, 2, 3, 0, 1, 2, 3], np.int32)
    x = np.ones(4)
    assert [orc.merge_path_search(d, ro) for d in (0, 4, 8, 12, 20)] == [(0, 0), (2, 2), (3, 5), (4, 8), (4, 8)]
    assert orc.thread_coords(3, ro).tolist() == [[0, 0], [2, 2], [3, 5], [4, 8]]
    for p in (1, 2, 3, 5, 12, 100):
        assert orc.merge_csrmv(ro, col, val, x, p).tolist() == [2, 0, 6, 16]
    assert orc.spmv_gold(ro, col, val, x).tolist() == [2, 0, 6, 16]


def test_device_spmv_docstring_lattice(orc):
    # cub/device/device_spmv.cuh:90-123: 3x3 lattice, values = x = 1 -> y = [2,3,2,3,4,3,2,3,2]
    ro = np.array([0, 2, 5, 7, 10, 14, 17, 19, 22, 24], np.int32)
    col = np.array([1, 3, 0, 2, 4, 1, 5, 0, 4, 6, 1, 3, 5, 7, 2, 4, 8, 3, 7, 4, 6, 8, 5, 7], np.int32)
    for dt in (np.float32, np.float64):
        y = orc.merge_csrmv(ro, col, np.ones(24, dt), np.ones(9, dt), 4)
        assert y.tolist() == [2, 3, 2, 3, 4, 3, 2, 3, 2]


def test_search_identity(orc):
    # x = #{r : row_end[r] + r < d}, past-the-end -> (rows, nnz)
    rng = np.random.default_rng(5)
    for _ in range(20):
        rows = int(rng.integers(1, 60))
        ro, _ = random_csr(rng, rows, 50, rng.uniform(0.2, 6), 0.3, 1)
        nnz = int(ro[-1])
        ends = ro[1:].astype(np.int64) + np.arange(rows)
        for d in range(0, rows + nnz + 5):
            x = min(int(np.sum(ends < d)), rows)
            want = (x, d - x) if d <= rows + nnz else (rows, nnz)
            assert orc.merge_path_search(d, ro) == want


# ---- golden vectors from the reference's own compiled code ---------------------------------------
def test_golden_coordinates(orc, golden_csrmv):
    g = golden_csrmv
    for c in range(int(g["num_cases"])):
        ro = g[f"c{c}_row_offsets"]
        got = np.array([orc.merge_path_search(int(d), ro) for d in g[f"c{c}_diags"]], np.int32)
        assert np.array_equal(got, g[f"c{c}_coords"]), f"case {c}"


@pytest.mark.parametrize("tag,dt", [("f64", np.float64), ("f32", np.float32)])
def test_golden_merge_csrmv_bit_exact(orc, golden_csrmv, tag, dt):
    # same thread count => same summation order => the restatement must match bit for bit
    g = golden_csrmv
    for c in range(int(g["num_cases"])):
        ro, col = g[f"c{c}_row_offsets"], g[f"c{c}_col"]
        val, x = g[f"c{c}_val_{tag}"], g[f"c{c}_x_{tag}"]
        assert np.array_equal(orc.spmv_gold(ro, col, val, x), g[f"c{c}_gold_{tag}"])
        for p in (1, 3, 8, 64):
            got = orc.merge_csrmv(ro, col, val, x, p)
            assert np.array_equal(got, g[f"c{c}_y_{tag}_p{p}"]), f"case {c} p={p}"


# ---- live against oracle/_ref (build container, and the GPU box where the .so travelled) ---------
def test_live_reference_random_structures(orc, ref):
    rng = np.random.default_rng(11)
    for it in range(25):
        rows = int(rng.integers(1, 400))
        cols = int(rng.integers(1, 300))
        ro, col = random_csr(rng, rows, cols, rng.uniform(0.1, 12), rng.uniform(0, 0.6), int(rng.integers(0, 3)))
        nnz = int(ro[-1])
        for d in rng.integers(0, rows + nnz + 40, 40):
            assert orc.merge_path_search(int(d), ro) == ref.merge_path_search(int(d), ro)
        for dt in (np.float32, np.float64):
            val = (0.5 + rng.random(nnz)).astype(dt)
            x = (0.5 + rng.random(cols)).astype(dt)
            assert np.array_equal(orc.spmv_gold(ro, col, val, x), ref.spmv_gold(ro, col, val, x))
            for p in (1, 2, 7, 256):
                a = orc.merge_csrmv(ro, col, val, x, p)
                b = ref.omp_merge_csrmv(ro, col, val, x, p)
                assert np.array_equal(a, b), f"iter {it} p={p} {dt}"


def test_alpha_beta_gold(orc, ref):
    rng = np.random.default_rng(3)
    ro, col = random_csr(rng, 50, 40, 4, 0.2)
    val, x, yin = rng.random(int(ro[-1])), rng.random(40), rng.random(50)
    a = orc.spmv_gold(ro, col, val, x, yin, alpha=1.7, beta=-0.3)
    b = ref.spmv_gold(ro, col, val, x, yin, alpha=1.7, beta=-0.3)
    assert np.array_equal(a, b)


def test_compare_results_rule(orc, ref):
    # utils.h:692-742: fp32-bit-pattern distance, FAIL iff sqrt(dist) > len; double narrows first
    rng = np.random.default_rng(9)
    for dt in (np.float32, np.float64):
        for n in (1, 4, 50):
            a = rng.random(n).astype(dt) + 1
            for bump in (0, 1e-7, 1e-5, 1e-3, 0.5):
                b = a.copy()
                b[n // 2] *= (1 + bump)
                assert (orc.compare_results(b, a) != 0) == (ref.compare_results(b, a) != 0)


def test_thread_count_invariance_exact_inputs(orc):
    # values = x = 1 -> y = row lengths for any p
    rng = np.random.default_rng(2)
    for _ in range(10):
        ro, col = random_csr(rng, int(rng.integers(1, 40)), 30, 3, 0.4)
        for p in (1, 2, 3, 5, 8, 17, 64, 255, 256, 1000):
            y = orc.merge_csrmv(ro, col, np.ones(int(ro[-1])), np.ones(30), p)
            assert np.array_equal(y, np.diff(ro).astype(np.float64))
