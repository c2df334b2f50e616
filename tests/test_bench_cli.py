"""bench.py contract pieces that run without a GPU: the reference arm's JSON line and the loud
failure of the product arm when CUDA is absent."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "cpu_uniform_16k", "--steps", "5", "--warmup", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "csr_spmv_gflops" and line["unit"] == "GFLOP/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "cpu_uniform_16k" and line["config"]["nnz"] == 524288


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_arm_refuses_without_cuda():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.parametrize("wname", ["cpu_uniform_16k", "powerlaw_2m"])
def test_bench_parity_check_accepts_the_oracle_and_rejects_a_wrong_row(orc, wname):
    """bench.check_parity (the check every bench line carries) on host tensors: the oracle's own result for a
    shard's rows passes; one corrupted row, a NaN, or a wrong all-ones result fails.  The reference it builds
    is an fp64 accumulation of the regenerated nonzeros, independent of any kernel."""
    import numpy as np

    import bench
    from merge_spmv_b200 import generators as gen
    from merge_spmv_b200 import sharded

    name, kind, dt, p, _ = bench.workload_spec(wname, 1)
    if wname == "powerlaw_2m":  # a small member of the same family
        p = dict(p, rows=3000, cols=3000, target_nnz=60000, max_row=2000)
    ro, cols, _ = bench.build_row_offsets(kind, p)
    dev = torch.device("cpu")
    col, val = bench.fill(kind, ro, cols, 0, int(ro[-1]), dt, "random", dev, p)
    x = gen.vector(cols, dt, "random")
    want = orc.merge_csrmv(ro.numpy(), col.numpy(), val.numpy(), x.numpy(), 4)
    world = 3
    for rank in range(world):
        shard = sharded.make_shard(ro.numpy(), cols, rank, world, lambda k0, k1: (col[k0:k1], val[k0:k1]), dev)
        y = torch.from_numpy(want[shard.x0:shard.x1].copy())
        ok = bench.check_parity(kind, ro, cols, dt, p, shard, x, y, "random", dev, chunk_nnz=5000)
        assert ok["ok"] and ok["rows_checked"] == shard.owned_rows, ok
        if shard.owned_rows:
            bad = y.clone()
            bad[shard.owned_rows // 2] *= 1.001
            assert not bench.check_parity(kind, ro, cols, dt, p, shard, x, bad, "random", dev)["ok"]
            bad = y.clone()
            bad[0] = float("nan")
            assert not bench.check_parity(kind, ro, cols, dt, p, shard, x, bad, "random", dev)["ok"]
            lens = torch.from_numpy(np.diff(ro.numpy())[shard.x0:shard.x1]).to(dt)
            assert bench.check_parity(kind, ro, cols, dt, p, shard, torch.ones(cols, dtype=dt), lens, "ones", dev)["ok"]
            assert not bench.check_parity(kind, ro, cols, dt, p, shard, torch.ones(cols, dtype=dt), lens + 1, "ones", dev)["ok"]
