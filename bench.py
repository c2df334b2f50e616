#!/usr/bin/env python
"""bench.py -- headline benchmark of the merge-based CsrMV path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--values ones|random]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU OmpMergeCsrmv, same config

A *step* is one y = A*x over the whole synthetic matrix through the C ABI (mspmv_csrmv_*), the
same unit the reference's timed loop repeats (gpu_spmv.cu:421-432).  Metric: GFLOP/s = 2*nnz/t
(gpu_spmv.cu:455,463).  At N=1 the default workload is BASELINE.json configs[1]: fp64, 1M x 1M,
64 nnz/row (826 MB of compulsory traffic per step, larger than the 126 MB L2, so no L2 flush is
needed between steps).  At N>1 the default is the same family grown with N (weak scaling: N*1M
rows, one merge-path shard per GPU, one exchange of N carries per step over NVLink); fixed-size
workloads (--workload powerlaw_20m ...) are sharded the same way and reported as "strong".

Prints ONE JSON line on rank 0.  Besides the contract keys it carries `parity` (this run's result against
an fp64 reference, every rank's rows), `roofline.secondary` (the SM->L2 request-port ceiling that bounds the
random-column workloads) and `extra_workloads` (BASELINE configs 3 and 4 at N=1, config 5 at every N).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Thread placement of the CPU legs (cpu_baseline, --impl reference): BASELINE.md section 4 prescribes
# OMP_PROC_BIND=spread OMP_PLACES=cores for the reference's OpenMP loop (its own KMP_AFFINITY is
# Intel-runtime-only).  libgomp reads these once, when it is first loaded -- i.e. before torch is imported.
os.environ.setdefault("OMP_PROC_BIND", "spread")
os.environ.setdefault("OMP_PLACES", "cores")
OMP_BINDING = f"OMP_PROC_BIND={os.environ['OMP_PROC_BIND']} OMP_PLACES={os.environ['OMP_PLACES']}"

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "csr_spmv_gflops"
UNIT = "GFLOP/s"


# --------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------
def workload_spec(name, n_gpus):
    """-> (kind, torch dtype, params dict, scaling, description)"""
    from merge_spmv_b200 import generators as gen

    if name == "default":
        name = "uniform_1m_64"
    kind, dt, p = gen.CONFIGS[name]
    p = dict(p)
    scaling = "weak"
    if name == "uniform_1m_64" and n_gpus > 1:
        # weak scaling: every GPU gets a config-2-sized shard (2^20 rows x 2^20 columns, 64 nnz/row) of
        # an (N * 2^20) x 2^20 matrix, so x -- replicated on every GPU -- keeps its N=1 size and the
        # per-GPU work is exactly the N=1 work.  (Growing the columns with N as well measures L2
        # capacity for x, not the path: see DESIGN.md section 6.)
        p["rows"] *= n_gpus
    elif n_gpus > 1:
        scaling = "strong"
    return name, kind, dt, p, scaling


def build_row_offsets(kind, p):
    from merge_spmv_b200 import generators as gen

    if kind in ("uniform", "uniform_local"):
        return gen.uniform_row_offsets(p["rows"], p["nnz_per_row"]), p["cols"], None
    if kind == "powerlaw":
        lengths, alpha = gen.powerlaw_row_lengths(p["rows"], min(p["max_row"], p["cols"]), p["target_nnz"])
        return gen._offsets_from_lengths(lengths), p["cols"], alpha
    return gen.banded_row_offsets(p["rows"], p["half_bandwidth"]), p["rows"], None


def fill(kind, row_offsets, cols, k0, k1, dt, values, device, p):
    from merge_spmv_b200 import generators as gen

    gkind = {"banded": "banded", "uniform_local": "local"}.get(kind, "stratified")
    return gen.fill_nonzeros(row_offsets, cols, k0, k1, kind=gkind, dtype=dt, values=values, device=device,
                             half_bandwidth=p.get("half_bandwidth", 3),
                             seed={"uniform": 0x5EED0001, "uniform_local": 0x5EED0001, "powerlaw": 0x5EED0003,
                                   "banded": 0x5EED0004}[kind])


def algorithmic_bytes(rows, cols, nnz, vb):
    # every array once
    return nnz * (vb + 4) + (rows + 1) * 4 + rows * vb + cols * vb


# --------------------------------------------------------------------------------------------------
# clocks (NVML, sampled in-process during the timed region)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        if self.ok:
            self._stop.set()
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvml unavailable"}
        reasons = sorted(r for r in self.reasons if r != "gpu_idle")
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_sm),
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# CPU baseline (oracle/_ref = the reference's own OmpMergeCsrmv; else the C port)
# --------------------------------------------------------------------------------------------------
def cpu_time(ro, col, val, x, budget_s):
    """Times the CPU merge CsrMV on this host in the reference's protocol.  Returns dict."""
    import oracle

    if oracle.Reference.available():
        impl, kind = oracle.Reference(), "reference"
        threads = min(impl.num_procs(), 256)  # cpu_spmv.cpp:302-303
        timer = impl.time_omp_merge_csrmv
    else:
        impl, kind = oracle.Oracle(), "port"
        threads = impl.num_procs()
        timer = impl.time_merge_csrmv
    nnz = int(ro[-1])
    ms1, _ = timer(ro, col, val, x, threads, 1)  # includes 4 untimed calls (cpu_spmv.cpp:380-392)
    iters = int(max(3, min(200, budget_s * 1000.0 / max(ms1, 1e-3))))
    ms, _ = timer(ro, col, val, x, threads, iters)
    return {"ms": ms, "gflops": 2.0 * nnz / ms / 1e6, "cores": threads, "kind": kind, "iters": iters, "nnz": nnz}


def host_sample(kind, dt, p, ro, cols, values, max_nnz):
    """The first rows of the workload holding at most max_nnz nonzeros, as host numpy arrays."""
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    nnz = int(ro[-1])
    rows = ro.numel() - 1
    if nnz > max_nnz:
        rows = int(torch.searchsorted(ro.to(torch.int64), torch.tensor(max_nnz), right=True)) - 1
        rows = max(rows, 1)
    ro_s = ro[: rows + 1].clone()
    k1 = int(ro_s[-1])
    col, val = fill(kind, ro, cols, 0, k1, dt, values, dev, p)
    return ro_s.numpy(), col.cpu().numpy(), val.cpu().numpy(), rows, k1


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU path, all host threads, same config and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from merge_spmv_b200 import generators as gen

    # the same N-scaled shape as the GPU arm; the CPU then times a bounded sample of it (leading rows)
    name, kind, dt, p, scaling = workload_spec(args.workload, max(args.gpus, 1))
    ro, cols, _ = build_row_offsets(kind, p)
    vb = 8 if dt == torch.float64 else 4
    full_rows, full_nnz = ro.numel() - 1, int(ro[-1])
    # bounded sample: cap the matrix so K+W steps finish within a few minutes on slow hosts
    budget_nnz = int(os.environ.get("MSPMV_REF_MAX_NNZ", 1 << 26))
    ro_s, col, val, rows, nnz = host_sample(kind, dt, p, ro, cols, args.values, budget_nnz)
    x = gen.vector(cols, dt, "ones" if args.values == "ones" else "random").numpy()
    import oracle

    if oracle.Reference.available():
        impl, kindname = oracle.Reference(), "reference"
        threads = min(impl.num_procs(), 256)
        step = lambda n: impl.time_omp_merge_csrmv(ro_s, col, val, x, threads, n)[0]
    else:
        impl, kindname = oracle.Oracle(), "port"
        threads = impl.num_procs()
        step = lambda n: impl.time_merge_csrmv(ro_s, col, val, x, threads, n)[0]
    # the reference harness itself does 4 untimed calls before timing (cpu_spmv.cpp:380-392)
    ms1 = step(1)
    steps = args.steps
    if ms1 * (steps + args.warmup) > 180e3:  # keep the whole run within a few minutes
        steps = max(3, int(180e3 / ms1) - args.warmup)
    if args.warmup > 4:
        step(args.warmup - 4)
    ms = step(steps)
    gflops = 2.0 * nnz / ms / 1e6
    sample = (f"first {rows} rows / {nnz} nnz of {name} ({'full workload' if nnz == int(ro[-1]) else 'bounded sample'}), "
              f"{steps} timed calls of OmpMergeCsrmv, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64" if vb == 8 else "f32", "data": "synthetic",
        "config": {"workload": name, "rows": full_rows, "cols": cols, "nnz": full_nnz, "values": args.values,
                   "sample_rows": rows, "sample_nnz": nnz, "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": gflops, "unit": UNIT, "cores": threads, "kind": kindname, "sample": sample,
                         "binding": OMP_BINDING},
        "e2e": {"value": gflops, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# parity inside the bench (outside the timed region): this rank's rows against an fp64 reference
# --------------------------------------------------------------------------------------------------
def check_parity(kind, ro, cols, dt, p, shard, x, y_owned, values, dev, chunk_nnz=1 << 25):
    """y_owned = global rows [shard.x0, shard.x1) as computed by the (sharded) CsrMV.  The reference
    regenerates those rows' nonzeros (the generators are counter-based, so any range can be rebuilt),
    accumulates val*x[col] per row in fp64 on the device (index_add), and is compared with the tolerances of
    tests/test_gpu_parity.py: exact for all-ones inputs, fp64 1e-10 relative, fp32
    max(1e-6, 4*sqrt(row length)*2^-24) relative."""
    x0, x1 = shard.x0, shard.x1
    n = x1 - x0
    if n == 0:
        return {"ok": True, "max_rel": 0.0, "rows_checked": 0, "tol": "n/a (rank owns no row)"}
    ro64 = ro.to(torch.int64)
    lens = (ro64[x0 + 1: x1 + 1] - ro64[x0: x1]).to(dev)
    if values == "ones":
        ok = bool(torch.equal(y_owned, lens.to(dt)))
        return {"ok": ok, "max_rel": 0.0 if ok else float("inf"), "rows_checked": n,
                "tol": "bit-exact: y == row lengths (values = x = 1, gpu_spmv.cu:521-525)"}
    ref = torch.zeros(n, dtype=torch.float64, device=dev)
    xd = x.double()
    r = x0
    while r < x1:  # whole rows per chunk, about chunk_nnz nonzeros each
        k0 = int(ro64[r])
        r_hi = int(torch.searchsorted(ro64, torch.tensor(k0 + chunk_nnz), right=True)) - 1
        r_hi = min(max(r_hi, r + 1), x1)
        k1 = int(ro64[r_hi])
        if k1 > k0:
            col, val = fill(kind, ro, cols, k0, k1, dt, values, dev, p)
            rl = (ro64[r + 1: r_hi + 1] - ro64[r: r_hi]).to(dev)
            rid = torch.repeat_interleave(torch.arange(r - x0, r_hi - x0, device=dev), rl)
            ref.index_add_(0, rid, val.double() * xd[col.long()])
            del col, val, rid
        r = r_hi
    err = (y_owned.double() - ref).abs()
    scale = ref.abs().clamp_min(1e-300)
    if dt == torch.float64:
        tol = torch.full_like(ref, 1e-10)
        tol_s = "fp64: |y - ref| <= 1e-10 |ref|"
    else:
        tol = torch.clamp(4 * torch.sqrt(lens.double()) * 2.0 ** -24, min=1e-6)
        tol_s = "fp32: |y - ref| <= max(1e-6, 4 sqrt(row length) 2^-24) |ref|"
    rel = err / scale
    ok = bool((rel <= tol).all()) and bool(torch.isfinite(y_owned).all())
    return {"ok": ok, "max_rel": float(rel.max()), "rows_checked": n, "tol": tol_s,
            "reference": "fp64 index_add of val*x[col] over the rank's rows, on the device"}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="default")
    ap.add_argument("--values", default="random", choices=["ones", "random"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--engine", default=None, choices=[None, "pipe", "auto"])
    ap.add_argument("--cols", type=int, default=None, help="override the column count (x length) of the workload")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay each step (memset node + the CsrMV kernel, plus the carry exchange at N>1) as one CUDA graph; auto = on")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE",
                    help="mspmv_set_option before the run (e.g. pipe_config=2, pipe_search=0); recorded in config.options")
    ap.add_argument("--exchange", default="p2p", choices=["nccl", "p2p"],
                    help="N>1: how the carries travel -- p2p (default): one kernel per rank storing into the peers' "
                         "symmetric memory over NVLink, fused with the fold (csrc/carry_exchange.cuh; 0.330 vs 0.409 ms "
                         "per step at 8 GPUs, profiles/mg_sweep_r02_n8.txt); nccl: one all_gather + fold kernel")
    ap.add_argument("--e2e-mode", default="auto", choices=["auto", "upload", "broadcast"],
                    help="N>1 end to end, how x reaches the GPUs -- upload: every rank copies it from its own pinned host "
                         "buffer over its own PCIe link; broadcast: it crosses PCIe once (rank 0) and NCCL broadcasts it "
                         "over NVLink; auto (default): both are timed and the faster one is reported as e2e, the other "
                         "under e2e.alternatives")
    ap.add_argument("--e2e-sequential", action="store_true",
                    help="N>1: run the end-to-end steps one after the other instead of the default three-stream "
                         "pipeline (H2D of x | broadcast + product | D2H of the y slice over three buffer slots)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra_workloads (BASELINE configs 3, 4 at N=1; config 5 at every N)")
    ap.add_argument("--gather-y", action="store_true",
                    help="N>1: include the all_gather of the y slices in every step (solver-style: the whole y "
                         "on every rank, ready to be the next x)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    import merge_spmv_b200 as ms
    from merge_spmv_b200 import generators as gen
    from merge_spmv_b200 import sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.engine:
        ms.lib().mspmv_set_engine(args.engine.encode())
    options = {}
    for kv in args.option:
        k, _, v = kv.partition("=")
        if ms.lib().mspmv_set_option(k.encode(), int(v)) != 0:
            raise SystemExit(f"unknown option {kv!r}")
        options[k] = int(v)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    L = ms.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(wname, steps, warmup, gather_y=False):
        """Build the workload, run warm-up + parity check + the timed region.  Returns a dict with the
        device-timed numbers plus the live objects (op, x, y, shard ...) the e2e / CPU legs re-use."""
        name, kind, dt, p, scaling = workload_spec(wname, world)
        if args.cols and wname == args.workload:
            p["cols"] = args.cols
        vb = 8 if dt == torch.float64 else 4
        ro, cols, alpha = build_row_offsets(kind, p)
        rows, nnz = ro.numel() - 1, int(ro[-1])
        ro_np = ro.numpy()
        x = gen.vector(cols, dt, "ones" if args.values == "ones" else "random", device=dev)
        shard = sharded.make_shard(ro_np, cols, rank, world,
                                   lambda k0, k1: fill(kind, ro, cols, k0, k1, dt, args.values, dev, p), dev)
        op = sharded.ShardedSpmv(shard, exchange=args.exchange if world > 1 else "nccl")
        use_graph = args.graph in ("on", "auto")
        c0 = L.mspmv_launch_count()
        op(x)  # one eager step: also counts this library's kernels per step (graph replays bypass the counter)
        launches_per_step = L.mspmv_launch_count() - c0
        gy = bool(gather_y and world > 1)
        if use_graph:
            step = op.capture(x, gather_y=gy)
        else:
            step = (lambda: op.matvec_full(x)) if gy else (lambda: op(x))

        # ---- warm-up, then parity OUTSIDE the timed region: this rank's rows against an fp64 reference
        for _ in range(warmup):
            y = step()
        torch.cuda.synchronize()
        y_owned = y[shard.x0:shard.x1] if gy else y
        if os.environ.get("MSPMV_BENCH_SKIP_PARITY") == "1":  # debugging aid only: the line then says so
            parity = {"ok": True, "max_rel": float("nan"), "rows_checked": 0, "tol": "SKIPPED (MSPMV_BENCH_SKIP_PARITY)"}
        else:
            parity = check_parity(kind, ro, cols, dt, p, shard, x, y_owned, args.values, dev)
        if world > 1:
            flag = torch.tensor([0 if parity["ok"] else 1], dtype=torch.int32, device=dev)
            worst = torch.tensor([parity["max_rel"]], dtype=torch.float64, device=dev)
            dist.all_reduce(flag)
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            parity["ok"] = bool(flag.item() == 0)
            parity["max_rel"] = float(worst.item())
            parity["ranks"] = world
        if not parity["ok"]:
            raise SystemExit(f"parity check failed on {name}: {parity}")

        # ---- timed region: exactly K steps, events on the launching stream, max over ranks ---------
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank)
        barrier()
        launches0 = L.mspmv_launch_count()
        trace = os.environ.get("MSPMV_BENCH_TRACE") == "1"  # debugging aid: one event per step, printed to stderr
        marks = []
        sync_token = torch.zeros(1, device=dev)
        with sampler:
            if world > 1:
                # device-side rendezvous right in front of the start event: the ranks leave the host barrier a
                # few ms apart, and without this the early rank's first step would time the late rank's arrival.
                # One more untimed step follows it (warm-up W+1): its carry exchange ends at the same moment on
                # every GPU, which aligns the start events to microseconds where the all-reduce leaves ~10 us.
                dist.all_reduce(sync_token)
                step()
            start.record()
            for _ in range(steps):
                step()
                if trace:
                    e = torch.cuda.Event(enable_timing=True)
                    e.record()
                    marks.append(e)
            stop.record()
            stop.synchronize()
        if trace:
            ts = [start.elapsed_time(e) for e in marks]
            print(f"[trace rank {rank}] {name}: per-step ms " +
                  " ".join(f"{b - a:.3f}" for a, b in zip([0.0] + ts[:-1], ts)), file=sys.stderr, flush=True)
        launches = L.mspmv_launch_count() - launches0
        if use_graph:
            launches = launches_per_step * steps  # replayed from the captured graph
        barrier()
        elapsed_ms = start.elapsed_time(stop)
        if world > 1:
            t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed_ms = float(t.item())
            lt = torch.tensor([launches], dtype=torch.int64, device=dev)
            dist.all_reduce(lt)
            launches = int(lt.item())
        ms_per_step = elapsed_ms / steps
        gflops = 2.0 * nnz / ms_per_step / 1e6

        # ---- roofline of the dominant kernel (per rank: its shard's compulsory bytes / step time) --
        shard_bytes = algorithmic_bytes(shard.local_rows, cols, shard.nnz, vb)
        achieved = shard_bytes / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{name}@{world}")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": "profiles/traffic.json: ncu dram__bytes_read+write of this kernel "
                                                          "build (static; tests/test_abi.py pins the SASS it was taken with)",
                    "peak_source": peak_src, "kernel": "spmv_pipe_kernel",
                    "algorithmic_bytes_per_launch": shard_bytes,
                    "frac_of_nominal_8000_gbs": achieved / 8000.0,  # BASELINE.md section 2 also asks for the nominal figure
                    "note": "one kernel per CsrMV (search, tiles and carry fold inside it); duration = whole step, CUDA events"}
        if kind in ("uniform", "powerlaw"):
            # Random columns: every nonzero costs one 32-byte sector request of x, and one SM sends ONE
            # request per clock to the L2 crossbar (ncu l1tex__m_l1tex2xbar_req_cycles_active; TMA bulk
            # loads share that port at one 128-byte request each).  That, not HBM, is the binding limit of
            # these workloads (profiles/gather_ceiling_r02.txt); reported beside the HBM roofline so frac
            # can be read against the right ceiling.
            sm_clk = (sampler.summary().get("sm_mhz") or 1965.0) * 1e6
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            requests = shard.nnz + (shard.nnz * (vb + 4) + shard.local_rows * 4) / 128.0
            r_rate = requests / (ms_per_step * 1e-3) / 1e9
            roofline["secondary"] = {"bound": "SM->L2 request port (1 request/clk/SM): one 32 B sector per x[col] gather "
                                              "+ one per 128 B of TMA tile data",
                                     "achieved": r_rate, "peak": sms * sm_clk / 1e9, "unit": "G requests/s",
                                     "frac": r_rate / (sms * sm_clk / 1e9),
                                     "peak_source": "148 SMs x SM clock during the run; counter evidence in "
                                                    "profiles/gather_ceiling_r02.txt"}
        return dict(name=name, kind=kind, dt=dt, p=p, scaling=scaling, vb=vb, ro=ro, ro_np=ro_np, cols=cols, rows=rows,
                    nnz=nnz, alpha=alpha, x=x, y=y, shard=shard, op=op, step=step, use_graph=use_graph, gather_y=gy,
                    ms_per_step=ms_per_step, gflops=gflops, launches=int(launches), roofline=roofline,
                    achieved=achieved, shard_bytes=shard_bytes, parity=parity, clocks=sampler.summary())

    m = measure(args.workload, args.steps, args.warmup, gather_y=args.gather_y)
    name, kind, dt, p, scaling, vb = m["name"], m["kind"], m["dt"], m["p"], m["scaling"], m["vb"]
    ro, ro_np, cols, rows, nnz, alpha = m["ro"], m["ro_np"], m["cols"], m["rows"], m["nnz"], m["alpha"]
    x, y, shard, op = m["x"], m["y"], m["shard"], m["op"]
    use_graph, gather_y = m["use_graph"], m["gather_y"]
    ms_per_step, gflops, launches, roofline = m["ms_per_step"], m["gflops"], m["launches"], m["roofline"]
    achieved, shard_bytes, parity_main = m["achieved"], m["shard_bytes"], m["parity"]

    # ---- e2e: host buffers through the public API, copies inside the timed region --------------
    e2e = None
    if not args.no_e2e:
        n_e2e = 50  # end-to-end steps (its own count: long enough that pipeline fill and first-use costs do not dominate)
        if world == 1:
            # session = upload A once (the driver's setup, gpu_spmv.cu:542-556), then per step:
            # x host->device, CsrMV, y device->host, pipelined over three streams
            ro_h, col_h, val_h = ro_np, shard.col.cpu().numpy(), shard.val.cpu().numpy()
            t0 = time.perf_counter()
            sess = ms.SpmvSession(ro_h, col_h, val_h, cols, device=local_rank)
            setup_ms = (time.perf_counter() - t0) * 1e3
            xs = torch.empty((n_e2e, cols), dtype=dt).pin_memory()
            ys = torch.empty((n_e2e, rows), dtype=dt).pin_memory()
            xs[:] = x.cpu()
            sess.apply_many(9, xs, ys)  # warm
            t0 = time.perf_counter()
            sess.apply_many(n_e2e, xs, ys)
            dt_s = time.perf_counter() - t0
            assert torch.equal(ys[n_e2e - 1].to(dev), y)
            t0 = time.perf_counter()
            for i in range(5):
                sess.apply(xs[i % n_e2e], ys[i % n_e2e])
            single_ms = (time.perf_counter() - t0) * 1e3 / 5
            sess.close()
            e2e = {"value": 2.0 * nnz * n_e2e / dt_s / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": cols * vb, "d2h_bytes_per_step": rows * vb, "steps": n_e2e,
                   "api": "mspmv_session_apply_many (pinned host x in, host y out, 3-stage pipeline)",
                   "unpipelined_ms_per_step": single_ms, "matrix_upload_ms": setup_ms,
                   "matrix_upload_bytes": int(nnz * (vb + 4) + (rows + 1) * 4)}
        else:
            # N > 1.  Every rank feeds its GPU from its own pinned host copy of x over its own PCIe link and
            # returns its slice of y to its own pinned host buffer (--e2e-mode broadcast: x crosses PCIe once, on rank
            # 0, and NCCL broadcasts it over NVLink).  Three streams over K slots:
            #   H2D(i+1) | product + carry exchange (i) | D2H(i-1)      (what mspmv_session_apply_many does at N=1)
            # --e2e-sequential: one step after the other.
            K = 3

            def e2e_variant(bcast):
                xh = x.cpu().pin_memory() if (rank == 0 or not bcast) else None
                s_h2d, s_comp, s_d2h = (torch.cuda.Stream(device=dev) for _ in range(3))
                xds = [torch.empty_like(x) for _ in range(K)]
                yds = [torch.empty(shard.owned_rows, dtype=dt, device=dev) for _ in range(K)]
                yhs = [torch.empty(shard.owned_rows, dtype=dt).pin_memory() for _ in range(K)]
                ev_x = [torch.cuda.Event() for _ in range(K)]
                ev_k = [torch.cuda.Event() for _ in range(K)]
                ev_y = [torch.cuda.Event() for _ in range(K)]

                # the device part of a slot (sharded CsrMV with its carry exchange, copy into the slot's y buffer) is
                # captured once per slot: the host then issues one graph launch per step.  With the p2p exchange the
                # graph holds no NCCL work (the NCCL paths -- --exchange nccl, --e2e-mode broadcast -- stay eager: replaying
                # graphs that contain collectives next to eager collectives costs milliseconds per switch)
                def device_part(sl):
                    if bcast:
                        dist.broadcast(xds[sl], src=0)
                    yds[sl].copy_(op(xds[sl]))

                slot_graphs = []
                if use_graph and not bcast and args.exchange == "p2p":
                    with torch.cuda.stream(s_comp):
                        for sl in range(K):
                            device_part(sl)  # warm-up outside capture
                    s_comp.synchronize()
                    for sl in range(K):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=s_comp):
                            device_part(sl)
                        slot_graphs.append(g)

                def run_pipelined(n):
                    for i in range(n):
                        sl = i % K
                        with torch.cuda.stream(s_h2d):
                            if i >= K:
                                s_h2d.wait_event(ev_k[sl])   # the product that read this x slot is done
                            if xh is not None:
                                xds[sl].copy_(xh, non_blocking=True)
                            ev_x[sl].record(s_h2d)
                        with torch.cuda.stream(s_comp):
                            s_comp.wait_event(ev_x[sl])
                            if i >= K:
                                s_comp.wait_event(ev_y[sl])  # the copy that drained this y slot is done
                            if slot_graphs:
                                slot_graphs[sl].replay()
                            else:
                                device_part(sl)
                            ev_k[sl].record(s_comp)
                        with torch.cuda.stream(s_d2h):
                            s_d2h.wait_event(ev_k[sl])
                            yhs[sl].copy_(yds[sl], non_blocking=True)
                            ev_y[sl].record(s_d2h)
                    for st in (s_h2d, s_comp, s_d2h):
                        st.synchronize()

                def run_sequential(n):
                    for i in range(n):
                        if xh is not None:
                            xds[0].copy_(xh, non_blocking=True)
                        if bcast:
                            dist.broadcast(xds[0], src=0)
                        yhs[0].copy_(op(xds[0]), non_blocking=True)
                    torch.cuda.synchronize()

                run = run_sequential if args.e2e_sequential else run_pipelined
                torch.cuda.synchronize()
                run(3 * K)  # warm
                barrier()
                t0 = time.perf_counter()
                run(n_e2e)
                barrier()
                dt_s = time.perf_counter() - t0
                last = yhs[0] if args.e2e_sequential else yhs[(n_e2e - 1) % K]
                assert torch.equal(last.to(dev), y if not gather_y else y[shard.x0:shard.x1])
                tt = torch.tensor([dt_s], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                return {"value": 2.0 * nnz * n_e2e / float(tt.item()) / 1e9, "unit": UNIT,
                       "h2d_bytes_per_step": cols * vb * (1 if bcast else world), "d2h_bytes_per_step": rows * vb,
                       "steps": n_e2e,
                       "api": ("ShardedSpmv: pinned host x on rank 0 -> one H2D -> NCCL broadcast over NVLink" if bcast else
                               "ShardedSpmv: every rank uploads x from its own pinned host copy over its own PCIe link")
                              + " -> sharded CsrMV + carry exchange -> every rank's y slice D2H to pinned host memory"
                              + (", sequential" if args.e2e_sequential else ", three-stream pipeline over 3 slots"),
                       "bytes_note": "totals over all ranks"}

            def e2e_mg_session():
                """The C-ABI multi-GPU operator (include/mergespmv.h section 5): rank 0 alone drives all N GPUs
                through mspmv_mg_session_apply_many -- x crosses PCIe once, peer copies over NVLink, every device
                returns its rows of y -- while the other ranks wait on the HOST (TCP store), their GPUs idle.
                Any failure is reported in the line instead of aborting the run."""
                store = dist.distributed_c10d._get_default_store()
                key = f"mg_e2e_done_{name}"
                res = None
                if rank == 0:
                    try:
                        np_dt = np.float64 if dt == torch.float64 else np.float32
                        col_h, val_h = np.empty(nnz, np.int32), np.empty(nnz, np_dt)
                        coords = sharded.partition(ro_np, world)
                        for g in range(world):  # the generators are counter-based: rebuild every shard's range here
                            k0, k1 = int(coords[g, 1]), int(coords[g + 1, 1])
                            if k1 > k0:
                                c, v = fill(kind, ro, cols, k0, k1, dt, args.values, dev, p)
                                col_h[k0:k1], val_h[k0:k1] = c.cpu().numpy(), v.cpu().numpy()
                                del c, v
                        t0 = time.perf_counter()
                        sess = ms.MultiGpuSpmvSession(ro_np, col_h, val_h, cols, list(range(world)))
                        setup_ms = (time.perf_counter() - t0) * 1e3
                        n_call, calls = 16, 3
                        xs = torch.empty((n_call, cols), dtype=dt).pin_memory()
                        ys = torch.empty((n_call, rows), dtype=dt).pin_memory()
                        xs[:] = x.cpu()
                        sess.apply_many(6, xs, ys)  # warm
                        t0 = time.perf_counter()
                        for _ in range(calls):
                            sess.apply_many(n_call, xs, ys)
                        dt_s = time.perf_counter() - t0
                        ok = bool(torch.equal(ys[n_call - 1][shard.x0:shard.x1].to(dev), y if not gather_y else y[shard.x0:shard.x1]))
                        dev_ms = sess.time_device(20)
                        sess.close()
                        res = {"value": 2.0 * nnz * n_call * calls / dt_s / 1e9, "unit": UNIT,
                               "h2d_bytes_per_step": cols * vb, "d2h_bytes_per_step": rows * vb, "steps": n_call * calls,
                               "api": "mspmv_mg_session_apply_many: ONE process (rank 0) drives all GPUs through the C ABI -- pinned "
                                      "host x -> one H2D -> peer copies over NVLink -> CsrMV + NVLink carry-exchange kernel per "
                                      "device -> every device's y rows D2H into one host vector; 3 slots x 3 streams per device",
                               "matches_sharded_result": ok, "device_ms_per_step": dev_ms, "matrix_upload_ms": setup_ms,
                               "bytes_note": "totals over all devices"}
                        if not ok:
                            res = {"error": "mg session result differs from the sharded result"}
                    except Exception as exc:  # never lose the line over the optional leg
                        res = {"error": repr(exc)[:300]}
                    finally:
                        store.set(key, "1")
                else:
                    store.wait([key])
                barrier()
                return res

            modes = {"upload": [False], "broadcast": [True], "auto": [False, True]}[args.e2e_mode]
            tried = [e2e_variant(b) for b in modes]
            if args.e2e_mode == "auto":
                mg = e2e_mg_session()
                if rank == 0 and mg is not None:
                    tried.append(mg if "value" in mg else {"value": -1.0, "api": "mspmv_mg_session_apply_many", "error": mg.get("error"),
                                                           "h2d_bytes_per_step": 0})
            e2e = max(tried, key=lambda r: r["value"])
            if len(tried) > 1:
                e2e["alternatives"] = [{k: r[k] for k in ("api", "value", "h2d_bytes_per_step", "error") if k in r}
                                       for r in tried if r is not e2e]

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        budget_nnz = int(os.environ.get("MSPMV_REF_MAX_NNZ", 1 << 26))
        ro_s, col_s, val_s, rows_s, nnz_s = host_sample(kind, dt, p, ro, cols, args.values, budget_nnz)
        c = cpu_time(ro_s, col_s, val_s, x.cpu().numpy(), budget_s=15.0)
        cpu = {"value": c["gflops"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"],
               "ms_per_step": c["ms"], "binding": OMP_BINDING,
               "sample": f"first {rows_s} rows / {nnz_s} nnz of {name} "
                         f"({'full workload' if nnz_s == nnz else 'bounded sample'}), {c['iters']} timed calls"}

    # ---- extra workloads: the other BASELINE.json configs, device-timed with parity, in the same line ----
    # N=1: config 3 (power-law fp32) and config 4 (banded fp64); every N: config 5 (20M x 20M power-law,
    # 1B nnz) sharded over the N GPUs -- "strong": the 8-GPU / 1-GPU ratio of its value is the north-star's >= 6x.
    extras = []
    if not args.no_extras and args.workload in ("default", "uniform_1m_64"):
        main_clocks = m["clocks"]
        step = None
        del m, op, shard, x, y
        torch.cuda.empty_cache()
        names = (["powerlaw_2m", "banded_10m"] if world == 1 else []) + ["powerlaw_20m"]
        for wn in names:
            k = 30 if wn == "powerlaw_20m" else 100
            e = measure(wn, k, 5)
            extras.append({"workload": e["name"], "dtype": "f64" if e["vb"] == 8 else "f32", "rows": e["rows"],
                           "cols": e["cols"], "nnz": e["nnz"], "scaling": "strong" if world > 1 else e["scaling"],
                           "steps": k, "ms_per_step": e["ms_per_step"], "value": e["gflops"], "unit": UNIT,
                           "roofline_frac": e["roofline"]["frac"], "hbm_gbs_algorithmic": e["achieved"] * world,
                           "request_port_frac": (e["roofline"].get("secondary") or {}).get("frac"),
                           "parity": e["parity"], "clocks": e["clocks"], "gpu_launches": e["launches"]})
            e["step"] = None
            del e
            torch.cuda.empty_cache()
    else:
        main_clocks = m["clocks"]

    if rank == 0:
        line = {
            "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64" if vb == 8 else "f32", "data": "synthetic",
            "config": {"workload": name, "rows": rows, "cols": cols, "nnz": nnz, "values": args.values,
                       "columns": {"banded": "banded", "uniform_local": "stratified within a 4096-column window around the diagonal"}.get(
                           kind, "stratified-uniform over all columns, sorted, distinct"),
                       "parallelism": f"merge-path shards x{world}" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (no flush)" if shard_bytes > 200e6 else "inputs fit L2",
                       "engine": "pipe", "cuda_graph": use_graph, "gather_y": gather_y, "options": options,
                       "carry_exchange": (args.exchange if world > 1 else None)},
            "roofline": roofline, "parity": parity_main, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": main_clocks, "extra_workloads": extras,
            "hbm_gbs_algorithmic": achieved * (1 if world == 1 else world),
        }
        if alpha is not None:
            line["config"]["powerlaw_alpha"] = alpha
        print(json.dumps(line), flush=True)
    if world > 1:
        # A captured graph holds NCCL work; tearing the communicator down under it can hang, so
        # release the graph first and leave without the (optional) collective teardown.
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        if use_graph:
            os._exit(0)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
