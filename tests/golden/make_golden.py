"""Regenerates tests/golden/*.npz from the reference's OWN code (oracle/_ref, compiled from
/root/reference by oracle/Makefile).  Runs only in the build container; the .npz files are
committed so the GPU box (which has no /root/reference) can check against them.

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def random_csr(rng, rows, cols, mean_len, empty_frac, long_rows):
    lens = rng.poisson(mean_len, rows)
    lens[rng.random(rows) < empty_frac] = 0
    for _ in range(long_rows):
        lens[rng.integers(rows)] = rng.integers(cols // 2, cols + 1)
    lens = np.minimum(lens, cols)
    ro = np.zeros(rows + 1, np.int32)
    ro[1:] = np.cumsum(lens)
    col = np.concatenate([np.sort(rng.choice(cols, l, replace=False)) for l in lens] + [np.zeros(0, int)])
    return ro, col.astype(np.int32)


def main():
    ref = oracle.Reference()
    rng = np.random.default_rng(20161113)  # SC'16

    # ---- (a)+(b): structures, merge-path coordinates, OmpMergeCsrmv outputs ------------------
    cases = {}
    shapes = [(1, 1, 1, 0.0, 0), (7, 5, 2, 0.3, 0), (40, 33, 3, 0.5, 1), (300, 257, 9, 0.1, 2),
              (1000, 64, 0.2, 0.8, 0), (513, 2048, 40, 0.0, 3), (2500, 2500, 7, 0.02, 0)]
    for ci, (rows, cols, mean_len, empty, longs) in enumerate(shapes):
        ro, col = random_csr(rng, rows, cols, mean_len, empty, longs)
        nnz = int(ro[-1])
        total = rows + nnz
        diags = np.unique(np.concatenate([rng.integers(0, total + 1, 64), [0, total, total + 1, total + 37]]))
        coords = np.array([ref.merge_path_search(int(d), ro) for d in diags], np.int32)
        cases[f"c{ci}_row_offsets"] = ro
        cases[f"c{ci}_col"] = col
        cases[f"c{ci}_cols"] = np.int32(cols)
        cases[f"c{ci}_diags"] = diags.astype(np.int32)
        cases[f"c{ci}_coords"] = coords
        for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
            val = (0.5 + rng.random(nnz)).astype(dt)
            x = (0.5 + rng.random(cols)).astype(dt)
            cases[f"c{ci}_val_{tag}"] = val
            cases[f"c{ci}_x_{tag}"] = x
            cases[f"c{ci}_gold_{tag}"] = ref.spmv_gold(ro, col, val, x)
            for p in (1, 3, 8, 64):
                cases[f"c{ci}_y_{tag}_p{p}"] = ref.omp_merge_csrmv(ro, col, val, x, num_threads=p)
    cases["num_cases"] = np.int32(len(shapes))
    np.savez_compressed(os.path.join(HERE, "merge_csrmv_ref.npz"), **cases)

    # ---- (c): the reference's generators + stats ----------------------------------------------
    gens = {}
    for kind, a, b in (("grid2d", 3, 0), ("grid2d", 6, 0), ("grid3d", 3, 0), ("grid3d", 4, 0),
                       ("wheel", 7, 0), ("wheel", 1, 0), ("dense", 4, 3), ("dense", 5, 8)):
        m = ref.build_matrix(kind, a, b)
        key = f"{kind}_{a}_{b}"
        for f in ("row_offsets", "col", "val", "stats"):
            gens[f"{key}_{f}"] = m[f]
        gens[f"{key}_dims"] = np.array([m["rows"], m["cols"], m["nnz"]], np.int32)
    np.savez_compressed(os.path.join(HERE, "generators_ref.npz"), **gens)

    # ---- (d): Matrix-Market reader quirks (sparse_matrix.h:217-380) ---------------------------
    mtx = {}
    files = {
        "general": "%%MatrixMarket matrix coordinate real general\n% comment\n4 5 6\n1 1 1.5\n3 2 -2\n1 4 3e1\n4 5 0.25\n3 2 7\n2 2 0\n",
        "symmetric": "%%MatrixMarket matrix coordinate real symmetric\n4 4 5\n1 1 2\n2 1 3\n3 2 4\n4 4 5\n4 1 -1\n",
        "skew": "%%MatrixMarket matrix coordinate real skew-symmetric\n3 3 2\n2 1 3\n3 1 -4\n",
        "pattern": "%%MatrixMarket matrix coordinate pattern general\n3 4 4\n1 2\n2 4\n3 1\n3 3\n",
        "array": "%%MatrixMarket matrix array real general\n3 2\n1\n2\n3\n4\n5\n6\n",
        "empty_rows": "%%MatrixMarket matrix coordinate real general\n6 6 3\n2 3 1\n2 1 2\n5 5 3\n",
    }
    with tempfile.TemporaryDirectory() as td:
        for name, text in files.items():
            path = os.path.join(td, name + ".mtx")
            with open(path, "w") as f:
                f.write(text)
            m = ref.build_matrix("market", path=path)
            mtx[f"{name}_text"] = np.frombuffer(text.encode(), dtype=np.uint8)
            for f_ in ("row_offsets", "col", "val"):
                mtx[f"{name}_{f_}"] = m[f_]
            mtx[f"{name}_dims"] = np.array([m["rows"], m["cols"], m["nnz"]], np.int32)
    mtx["names"] = np.array(sorted(files))
    np.savez_compressed(os.path.join(HERE, "market_ref.npz"), **mtx)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
