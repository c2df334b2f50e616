"""The C-ABI library loads and exports every symbol include/mergespmv.h declares; host-only
entry points agree with the oracle.  CPU only -- no compute call is made without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, random_csr

import merge_spmv_b200 as ms
from merge_spmv_b200 import _lib, sharded


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mergespmv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mspmv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 20
    L = C.CDLL(ms.lib_path())
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/mergespmv.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)


def test_version_and_error_string():
    L = ms.lib()
    assert L.mspmv_version() == 2  # MSPMV_VERSION_MAJOR * 100 + MINOR (0.2: round 2, pipe engine + mg session)
    assert b"invalid argument" in L.mspmv_error_string(1)
    assert L.mspmv_set_engine(b"bogus") == 1
    assert L.mspmv_set_engine(b"auto") == 0


def test_no_cpu_fallback():
    import torch
    ro = torch.tensor([0, 1], dtype=torch.int32)
    with pytest.raises(ms.MergeSpmvError):
        ms.csrmv(ro, torch.zeros(1, dtype=torch.int32), torch.ones(1), torch.ones(1))


def test_missing_library_fails_loudly():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import merge_spmv_b200 as ms\n"
            "try:\n    ms.lib()\nexcept ms.MergeSpmvError as e:\n    print('LOUD', e)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       env=dict(os.environ, MSPMV_LIB="/nonexistent/libmergespmv.so"), timeout=120)
    assert "LOUD" in r.stdout and "no CPU fallback" in r.stdout, r.stdout + r.stderr


def test_host_merge_path_search_matches_oracle(orc):
    rng = np.random.default_rng(17)
    L = ms.lib()
    for _ in range(30):
        rows = int(rng.integers(1, 300))
        ro, _ = random_csr(rng, rows, 64, rng.uniform(0.1, 9), rng.uniform(0, 0.5), 1)
        nnz = int(ro[-1])
        for d in rng.integers(0, rows + nnz + 30, 50):
            x, y = C.c_int(), C.c_int()
            L.mspmv_host_merge_path_search(ro.ctypes.data_as(C.c_void_p), rows, nnz, int(d), C.byref(x), C.byref(y))
            assert (x.value, y.value) == orc.merge_path_search(int(d), ro)


def test_shard_partition_equals_cpu_thread_coords(orc):
    # shard g of p == thread g of OmpMergeCsrmv with p threads (cpu_spmv.cpp:311-321)
    rng = np.random.default_rng(23)
    for _ in range(20):
        rows = int(rng.integers(1, 500))
        ro, _ = random_csr(rng, rows, 100, rng.uniform(0.1, 20), rng.uniform(0, 0.5), 2)
        for p in (1, 2, 3, 4, 8, 16):
            assert np.array_equal(sharded.partition(ro, p), orc.thread_coords(p, ro))


def test_shard_row_offsets_cover_matrix():
    rng = np.random.default_rng(29)
    rows = 257
    ro, _ = random_csr(rng, rows, 90, 6, 0.2, 2)
    for p in (1, 2, 5, 8):
        coords = sharded.partition(ro, p)
        total = 0
        for g in range(p):
            lro = sharded.local_row_offsets(ro, coords, g)
            (x0, y0), (x1, y1) = coords[g], coords[g + 1]
            assert lro[0] == 0 and lro[-1] == y1 - y0 and np.all(np.diff(lro) >= 0)
            assert lro.size == x1 - x0 + 2
            total += lro[-1]
        assert total == ro[-1]


def test_default_kernels_are_the_measured_ones():
    """The round-2 numbers in profiles/ and DESIGN.md section 5 were measured with a specific build.  Every
    kernel of that build must still be in libmergespmv.so with the same SASS (operand order inside
    an instruction aside), so that work done without GPU time cannot silently change what the
    bench runs by default.  After a kernel is changed on purpose AND re-measured, regenerate the
    file: python tools/sass_fingerprint.py > profiles/sass_fingerprint_rNN.json."""
    import importlib.util
    import json
    import shutil

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    spec = importlib.util.spec_from_file_location("sass_fingerprint", os.path.join(ROOT, "tools", "sass_fingerprint.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    measured = json.load(open(os.path.join(ROOT, "profiles", "sass_fingerprint_r02.json")))
    now = mod.fingerprint(os.path.join(ROOT, "merge-spmv_b200", "libmergespmv.so"))
    changed = [k for k, v in measured.items() if now.get(k) != v]
    assert not changed, f"kernels differ from the measured build: {changed}"
    assert all(k in now for k in measured), "a measured kernel is missing from the library"


def test_sass_tools_smoke():
    """tools/sass_lines.py maps the SASS of a kernel back to source lines (the library is built with
    -lineinfo): the shipped pipe kernel must show TMA bulk copies and attribute instructions to
    spmv_pipe.cuh."""
    import shutil
    import subprocess
    import sys

    if shutil.which("nvdisasm") is None or shutil.which("cuobjdump") is None:
        pytest.skip("CUDA binary utilities not on PATH")
    tool = os.path.join(ROOT, "tools", "sass_lines.py")
    by_file = subprocess.run([sys.executable, tool, "spmv_pipe_kernelINS_7PipeCfgId", "--by-file"], capture_output=True, text=True,
                             check=True).stdout
    assert "spmv_pipe.cuh" in by_file and "tma_stage.cuh" in by_file
    ops = subprocess.run([sys.executable, tool, "spmv_pipe_kernelINS_7PipeCfgId", "--ops"], capture_output=True, text=True,
                         check=True).stdout
    assert "UBLKCP" in ops and "SYNCS" in ops, "TMA bulk copy / mbarrier instructions missing from the pipe kernel"
