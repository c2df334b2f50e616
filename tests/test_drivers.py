"""The C++ driver surface (merge-spmv_b200/host): Matrix-Market reader and generators pinned to the
golden CSR the reference's own code produced, the CLI/CSV contract, and (GPU) the gpu_spmv binary's
self-check.  The cpu_spmv driver is a separate CPU tool and runs here without a GPU."""
import os
import subprocess
import sys

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


def test_matrix_market_round_trip_and_corpus_script(tmp_path):
    """--writemtx output re-read through --mtx gives the same CSR (so synthetic configs can be fed to
    the reference's own binaries), and eval_csrmv.sh prints the reference's CSV header + one line
    per .mtx file (eval_csrmv.sh:8,14-17)."""
    a, b = tmp_path / "a.bin", tmp_path / "b.bin"
    corpus = tmp_path / "corpus"
    corpus.mkdir()
    mtx = corpus / "u.mtx"
    assert run([CPU, "--uniform=5", "--rows=300", "--values=random", "--quiet", "--i=1", f"--dumpcsr={a}",
                f"--writemtx={mtx}"]).returncode == 0
    assert run([CPU, f"--mtx={mtx}", "--quiet", "--i=1", f"--dumpcsr={b}"]).returncode == 0
    ra, rb = read_dump(a), read_dump(b)
    assert ra[:3] == rb[:3]
    for x, y in zip(ra[3:], rb[3:]):
        assert np.array_equal(x, y)
    assert run([CPU, "--grid2d=12", "--quiet", "--i=1", f"--writemtx={corpus / 'g.mtx'}"]).returncode == 0
    r = run([os.path.join(PKG, "eval_csrmv.sh"), str(corpus), "cpu_spmv"])
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    assert lines[0].startswith("file, num_rows, num_cols, num_nonzeros, row_length_mean")
    assert len(lines) == 3 and all("Merge CsrMV" in l for l in lines[1:])


def _reference_driver_stdout(args):
    """stdout of the reference's own cpu_spmv main() (compiled into oracle/_ref), in a subprocess."""
    import oracle
    code = ("import ctypes as C, sys; L = C.CDLL(%r); a = [b'cpu_spmv'] + [s.encode() for s in sys.argv[1:]]; "
            "argv = (C.c_char_p * len(a))(*a); L.ref_cpu_spmv_main(len(a), argv); "
            "C.CDLL(None).fflush(None)" % oracle.Reference.path())
    return subprocess.run([sys.executable, "-c", code] + args, capture_output=True, text=True, timeout=300).stdout


@pytest.mark.parametrize("flags", [["--grid2d=40"], ["--grid3d=12"], ["--dense=16", "--fp32"]])
def test_csv_contract_matches_reference_driver(ref, flags):
    """--quiet line of our cpu_spmv vs the reference's own main(): same label and the same seven
    statistics fields, character for character; the method name "Merge CsrMV" with four numbers.
    (The reference also prints an "MKL CsrMV" group first, cpu_spmv.cpp:649-652; MKL is out of scope.)"""
    theirs = [f.strip() for f in _reference_driver_stdout(flags + ["--quiet", "--i=2"]).strip().rstrip(",").split(",")]
    ours = [f.strip() for f in run([CPU] + flags + ["--quiet", "--i=2"]).stdout.strip().rstrip(",").split(",")]
    assert ours[:8] == theirs[:8], (ours[:8], theirs[:8])
    assert "MKL CsrMV" in theirs and theirs[theirs.index("Merge CsrMV") - 5] == "MKL CsrMV"
    i, j = ours.index("Merge CsrMV"), theirs.index("Merge CsrMV")
    assert len(ours[i + 1:]) == len(theirs[j + 1:]) == 4
    for a in ours[i + 1:]:
        float(a)


def test_human_output_matches_reference_driver(ref):
    """Non-quiet mode: the statistics block, the histogram and the perf line grammar are the
    reference's (sparse_matrix.h:75-89,919-956; cpu_spmv.cpp:515-520)."""
    import re
    flags = ["--grid3d=10", "--i=2", "--threads=2"]
    theirs = _reference_driver_stdout(flags).splitlines()
    ours = run([CPU] + flags).stdout.splitlines()
    keep = lambda ls: [l for l in ls if re.search(r"num_rows:|num_cols:|num_nonzeros:|row_length_|CSR matrix \(|Degree 1e|Using 2 threads", l)]
    assert keep(ours) == keep(theirs) and len(keep(ours)) >= 11
    perf = re.compile(r"^fp64: \d+\.\d{4} setup ms, \d+\.\d{4} avg ms, \d+\.\d{5} gflops, \d+\.\d{3} effective GB/s$")
    assert any(perf.match(l) for l in ours) and any(perf.match(l) for l in theirs)
    assert any(l.strip() == "PASS" for l in ours)


def test_v2_matrix_display_matches_reference_driver(ref):
    """--v2 prints the input matrix (cpu_spmv.cpp:604 -> CsrMatrix::Display, sparse_matrix.h:962-975):
    the same lines as the reference's own main(), character for character."""
    flags = ["--grid2d=4", "--i=1", "--v2", "--threads=1"]
    grab = lambda ls: [l for l in ls if l.startswith("Input Matrix") or " [@" in l]
    theirs, ours = grab(_reference_driver_stdout(flags).splitlines()), grab(run([CPU] + flags).stdout.splitlines())
    assert ours == theirs and len(ours) == 17


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
    # --v: reference and computed vectors (utils.h:785-798); --v2: the input matrix; --gpus: the multi-GPU session
    r = run([GPU, "--grid2d=5", "--i=2", "--v", "--v2"])
    assert "Reference:" in r.stdout and "Computed:" in r.stdout and "Input Matrix (25 vertices, 80 nonzeros):" in r.stdout
    import torch
    n = torch.cuda.device_count()
    if n >= 2:
        r = run([GPU, "--uniform=64", f"--rows={65536 * n}", "--cols=65536", "--values=random", "--randx", "--i=20", f"--gpus={n}"])
        assert r.returncode == 0 and "FAIL" not in r.stdout and r.stdout.count("PASS") == 2 and "MISMATCH" not in r.stdout, r.stdout
    r = run([GPU, "--grid2d=300", "--quiet", "--i=10"])
    fields = [f.strip() for f in r.stdout.strip().rstrip(",").split(",")]
    # label, 7 stats, device, fp64, method, 4 perf numbers (gpu_spmv.cu:532-534,467-471)
    assert fields[0] == "grid2d_300" and fields[9] == "fp64" and fields[10] == "Merge-based CsrMV" and len(fields) == 15
