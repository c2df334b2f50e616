#!/usr/bin/env python
"""Per-phase SM cycles of the pipe kernel's consumer loop (thread 0 of every block), from the tuning build
merge-spmv_b200/variants/libmergespmv_prof.so (make -C merge-spmv_b200 variants/libmergespmv_prof.so):

    MSPMV_LIB=merge-spmv_b200/variants/libmergespmv_prof.so python tools/phase_profile.py [workloads]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import merge_spmv_b200 as ms  # noqa: E402
from merge_spmv_b200 import generators as gen, sharded  # noqa: E402

NAMES = ["issue (smem reads, cols, gathers)", "walk (waits for x)", "scan", "wait next tile + P1", "barrier", "Y + release",
         None, "between tiles"]
L = ms.lib()
L.mspmv_debug_profile.argtypes = [C.c_void_p, C.c_int]
dev = torch.device("cuda", 0)
for wname in (sys.argv[1:] or ["uniform_1m_64", "powerlaw_2m", "banded_10m"]):
    name, kind, dt, p, _ = bench.workload_spec(wname, 1)
    ro, cols, _ = bench.build_row_offsets(kind, p)
    x = gen.vector(cols, dt, "random", device=dev)
    shard = sharded.make_shard(ro.numpy(), cols, 0, 1, lambda k0, k1: bench.fill(kind, ro, cols, k0, k1, dt, "random", dev, p), dev)
    op = sharded.ShardedSpmv(shard)
    for _ in range(3):
        op(x)
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 8)()
    L.mspmv_debug_profile(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        op(x)
    e1.record()
    e1.synchronize()
    L.mspmv_debug_profile(out, 0)
    tiles = out[6]
    total = sum(out[k] for k in range(8) if k != 6)
    print(f"== {name}: {e0.elapsed_time(e1) / n:.4f} ms/step (profiling build), {tiles // n} tiles/step, "
          f"{total / max(tiles, 1):.0f} cycles per tile per block")
    for k, nm in enumerate(NAMES):
        if nm:
            print(f"   {nm:36s} {out[k] / max(tiles, 1):8.0f} cycles/tile  {100.0 * out[k] / max(total, 1):5.1f} %")
    del op, shard, x
    torch.cuda.empty_cache()
