"""The C++ driver surface (merge-spmv_b200/host): Matrix-Market reader and generators pinned to the
golden CSR the reference's own code produced, the CLI/CSV contract, and (GPU) the gpu_spmv binary's
self-check.  The cpu_spmv driver is a separate CPU tool and runs here without a GPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

PKG = os.path.join(ROOT, "merge-spmv_b200")
CPU = os.path.join(PKG, "cpu_spmv")
GPU = os.path.join(PKG, "gpu_spmv")


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


def read_dump(path):
    raw = np.fromfile(path, dtype=np.int32, count=3)
    rows, cols, nnz = map(int, raw)
    with open(path, "rb") as f:
        f.seek(12)
        ro = np.fromfile(f, np.int32, rows + 1)
        col = np.fromfile(f, np.int32, nnz)
        val = np.fromfile(f, np.float64, nnz)
    return rows, cols, nnz, ro, col, val


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(os.path.join(PKG, "bin", "_cpu_spmv_driver")):
        subprocess.run(["make", "-C", PKG, "-s"], check=True)


def test_matrix_market_reader_matches_reference(tmp_path):
    g = np.load(os.path.join(GOLDEN, "market_ref.npz"))
    for name in g["names"]:
        name = str(name)
        mtx = tmp_path / f"{name}.mtx"
        mtx.write_bytes(g[f"{name}_text"].tobytes())
        dump = tmp_path / f"{name}.bin"
        r = run([CPU, f"--mtx={mtx}", "--quiet", "--i=1", f"--dumpcsr={dump}"])
        assert r.returncode == 0, r.stderr
        rows, cols, nnz, ro, col, val = read_dump(dump)
        assert [rows, cols, nnz] == g[f"{name}_dims"].tolist(), name
        assert np.array_equal(ro, g[f"{name}_row_offsets"]), name
        assert np.array_equal(col, g[f"{name}_col"]), name
        assert np.array_equal(val, g[f"{name}_val"]), name


def test_generators_match_reference(tmp_path):
    g = np.load(os.path.join(GOLDEN, "generators_ref.npz"))
    for flag, key in (("--grid2d=6", "grid2d_6_0"), ("--grid3d=4", "grid3d_4_0"), ("--wheel=7", "wheel_7_0")):
        dump = tmp_path / "m.bin"
        r = run([CPU, flag, "--quiet", "--i=1", f"--dumpcsr={dump}"])
        assert r.returncode == 0, r.stderr
        rows, cols, nnz, ro, col, val = read_dump(dump)
        assert [rows, cols, nnz] == g[f"{key}_dims"].tolist()
        assert np.array_equal(ro, g[f"{key}_row_offsets"]) and np.array_equal(col, g[f"{key}_col"])
        # quiet CSV: label, 3 ints, 4 stats, method, 4 perf numbers (README.md:151, eval_csrmv.sh:8)
        fields = [f.strip() for f in r.stdout.strip().rstrip(",").split(",")]
        assert fields[0] == flag[2:].replace("=", "_") and fields[8] == "Merge CsrMV" and len(fields) == 13
        stats = g[f"{key}_stats"]
        assert abs(float(fields[4]) - stats[0]) < 1e-4 and abs(float(fields[5]) - stats[1]) < 1e-4


def test_synthetic_families_match_python_generators(tmp_path):
    import torch
    from merge_spmv_b200 import generators as gen

    dump = tmp_path / "u.bin"
    assert run([CPU, "--uniform=32", "--rows=4096", "--values=random", "--quiet", "--i=1", f"--dumpcsr={dump}"]).returncode == 0
    rows, cols, nnz, ro, col, val = read_dump(dump)
    m = gen.uniform(4096, 4096, 32, values="random")
    assert np.array_equal(ro, m.row_offsets.numpy()) and np.array_equal(col, m.col.numpy())
    assert np.array_equal(val, m.val.numpy())
    assert run([CPU, "--banded=3", "--rows=5000", "--quiet", "--i=1", f"--dumpcsr={dump}"]).returncode == 0
    rows, cols, nnz, ro, col, val = read_dump(dump)
    b = gen.banded(5000, 3)
    assert np.array_equal(ro, b.row_offsets.numpy()) and np.array_equal(col, b.col.numpy())
    assert run([CPU, "--powerlaw=500", "--rows=3000", "--nnz=60000", "--fp32", "--quiet", "--i=1", f"--dumpcsr={dump}"]).returncode == 0
    rows, cols, nnz, ro, col, val = read_dump(dump)
    p = gen.powerlaw(3000, 3000, 500, 60000)
    assert np.array_equal(ro, p.row_offsets.numpy()) and np.array_equal(col, p.col.numpy())


def test_cpu_driver_pass_and_threads():
    r = run([CPU, "--grid3d=30", "--i=5", "--threads=3"])
    assert r.returncode == 0 and "PASS" in r.stdout and "Using 3 threads" in r.stdout
    assert "row_length_mean" in r.stdout and "Degree 1e0" in r.stdout
    r = run([CPU, "--uniform=32", "--rows=16384", "--i=5"])  # BASELINE.json configs[0]
    assert "PASS" in r.stdout and "gflops" in r.stdout


@pytest.mark.gpu
def test_gpu_driver_self_check():
    for flags in (["--grid2d=500"], ["--wheel=100000", "--fp32"], ["--dense=64", "--size=1048576"],
                  ["--uniform=64", "--rows=65536", "--values=random", "--randx"],
                  ["--banded=3", "--rows=200000", "--alpha=2.0", "--beta=0.5"],
                  ["--powerlaw=20000", "--rows=50000", "--nnz=3000000", "--fp32", "--cusparse"]):
        r = run([GPU, "--i=20"] + flags)
        assert r.returncode == 0, r.stderr
        assert "FAIL" not in r.stdout and r.stdout.count("PASS") >= 1, r.stdout
        assert "Merge-based CsrMV" in r.stdout and "effective GB/s" in r.stdout
    r = run([GPU, "--grid2d=300", "--quiet", "--i=10"])
    fields = [f.strip() for f in r.stdout.strip().rstrip(",").split(",")]
    # label, 7 stats, device, fp64, method, 4 perf numbers (gpu_spmv.cu:532-534,467-471)
    assert fields[0] == "grid2d_300" and fields[9] == "fp64" and fields[10] == "Merge-based CsrMV" and len(fields) == 15
