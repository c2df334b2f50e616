#!/usr/bin/env bash
# gather-bound workloads: pipe variants x shared-memory budget (the L1 left over bounds the misses in flight)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"; LOG="$OUT/pipe_sweep2.txt"; : > "$LOG"
V=merge-spmv_b200/variants
for KB in ${SMEM_KBS:-112 128 144}; do
  for B in ${VARIANTS:-p_c1 p_ipt7_c1 p_ipt11_c1 p_nw2_c1 p_nw2_ipt13 p_nw2_ipt11_c1 p_nw3_c1}; do
    MSPMV_PIPE_SMEM_KB=$KB MSPMV_LIB=$V/libmergespmv_$B.so timeout 600 python tools/sweep_lib.py --label "$B@${KB}KB" --steps 200 \
        --options "engine=pipe" --workloads ${WL:-uniform_1m_64,powerlaw_2m} 2>&1 | grep -E "\|" | tee -a "$LOG"
  done
done
echo done | tee -a "$LOG"
