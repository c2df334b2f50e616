#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + hottest source lines.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [--lines N]"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_config_size", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum",
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "branch_resolving", "membar",
               "mio_throttle", "lg_throttle", "math_pipe_throttle", "no_instruction", "sleeping",
               "not_selected", "dispatch_stall", "drain", "tex_throttle", "imc_miss"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")][:90])
        for m in METRICS:
            if m in hdr:
                print(f"  {m:70s} {r[hdr.index(m)]} {units[hdr.index(m)]}")
        st = []
        for s in STALL_NAMES:
            m = STALLS % s
            if m in hdr:
                st.append((float(r[hdr.index(m)] or 0), s))
        print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    src = run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    rows = list(csv.reader(io.StringIO(src)))
    agg = {}      # (file, line) -> [samples, warp_inst, thread_inst, text]
    fname, hdr = "?", None
    tot_s = tot_i = 0.0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            c_s = hdr.index("# Samples")
            c_i = hdr.index("Instructions Executed")
            c_t = hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        try:
            sm_, wi, ti = float(r[c_s] or 0), float(r[c_i] or 0), float(r[c_t] or 0)
        except ValueError:
            continue
        a = agg.setdefault((fname, int(r[0])), [0.0, 0.0, 0.0, r[1].strip()[:100]])
        a[0] += sm_
        a[1] += wi
        a[2] += ti
        tot_s += sm_
        tot_i += wi
    print(f"== hottest source lines (stall samples {tot_s:.0f}, warp instructions {tot_i:.0f})")
    print("   samples%  inst%   file:line  source")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:nlines]:
        print(f"  {100*a[0]/max(tot_s,1):6.1f}% {100*a[1]/max(tot_i,1):6.1f}%  {f}:{ln}: {a[3]}")
    print("== most executed source lines")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:nlines]:
        print(f"  {100*a[1]/max(tot_i,1):6.1f}% (samples {100*a[0]/max(tot_s,1):5.1f}%)  {f}:{ln}: {a[3]}")


if __name__ == "__main__":
    main()
