#!/usr/bin/env python
"""Time one build of libmergespmv.so (MSPMV_LIB, default: the shipped one) on several workloads and
runtime options in ONE process: the matrix of a workload is generated once, every option is timed on
it with the CUDA-graph replay bench.py uses.  One line per (workload, option):

    label | workload | option | ms per step | GFLOP/s | algorithmic GB/s | fraction of the HBM peak

    MSPMV_LIB=merge-spmv_b200/variants/libmergespmv_ipt8_12.so python tools/sweep_lib.py --label ipt8_12
    MSPMV_TILE_CARVEOUT=56 python tools/sweep_lib.py --label carve56 --workloads uniform_1m_64,powerlaw_2m

Compile-time switches and the shared-memory carve-out are fixed per process (the library reads the
carve-out once per kernel); pipe_config / pipe_search are switched at run time.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--label", default="shipped")
    ap.add_argument("--workloads", default="uniform_1m_64,powerlaw_2m,banded_10m,uniform_1m_64_local")
    ap.add_argument("--options", default="engine=pipe;engine=pipe,pipe_search=0",
                    help="semicolon-separated option sets, each a comma-separated list of name=value")
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--values", default="random")
    args = ap.parse_args()

    import bench
    import merge_spmv_b200 as ms
    from merge_spmv_b200 import generators as gen
    from merge_spmv_b200 import sharded

    assert torch.cuda.is_available(), "needs a CUDA device"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = ms.lib()
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    print(f"# {args.label}: {ms.lib_path()} carve-out={os.environ.get('MSPMV_TILE_CARVEOUT', 'default')} peak={peak} GB/s",
          flush=True)
    for wname in args.workloads.split(","):
        name, kind, dt, p, _ = bench.workload_spec(wname, 1)
        vb = 8 if dt == torch.float64 else 4
        ro, cols, _ = bench.build_row_offsets(kind, p)
        rows, nnz = ro.numel() - 1, int(ro[-1])
        x = gen.vector(cols, dt, "ones" if args.values == "ones" else "random", device=dev)
        shard = sharded.make_shard(ro.numpy(), cols, 0, 1,
                                   lambda k0, k1: bench.fill(kind, ro, cols, k0, k1, dt, args.values, dev, p), dev)
        op = sharded.ShardedSpmv(shard)
        nbytes = bench.algorithmic_bytes(rows, cols, nnz, vb)
        reference = None
        for optset in args.options.split(";"):
            opts = dict(kv.split("=") for kv in optset.split(",") if kv)
            for k, v in opts.items():
                if k == "engine":
                    assert L.mspmv_set_engine(v.encode()) == 0, f"unknown engine {v}"
                else:
                    assert L.mspmv_set_option(k.encode(), int(v)) == 0, f"unknown option {k}"
            step = op.capture(x)
            for _ in range(args.warmup):
                y = step()
            torch.cuda.synchronize()
            if reference is None:
                reference = y.clone()
            same = bool(torch.equal(y, reference))
            if not same:  # engines differ in summation order: report the distance instead
                rel = ((y - reference).abs() / reference.abs().clamp_min(1e-30)).max().item()
                same = f"max rel diff {rel:.2e}"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            e1.synchronize()
            ms_step = e0.elapsed_time(e1) / args.steps
            gbs = nbytes / (ms_step * 1e-3) / 1e9
            print(f"{args.label:18s} | {name:20s} | {optset:28s} | {ms_step:8.4f} ms | {2.0 * nnz / ms_step / 1e6:8.1f} GFLOP/s | "
                  f"{gbs:7.0f} GB/s | {gbs / peak:5.3f} | {'same bits' if same is True else same}", flush=True)
            for k in opts:
                if k == "engine":
                    L.mspmv_set_engine(b"auto")
                else:
                    L.mspmv_set_option(k.encode(), -1)
            del step
        del op, shard, x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
