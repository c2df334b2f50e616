#!/usr/bin/env bash
# A/B of library builds on the four single-GPU workloads: LIBS="a.so b.so" bash tools/ab.sh
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"; LOG="$OUT/ab.txt"; : > "$LOG"
for L in ${LIBS}; do
    MSPMV_LIB=$L timeout 600 python tools/sweep_lib.py --label "$(basename $L .so)" --steps 300 --options "engine=pipe" \
        --workloads ${WL:-uniform_1m_64,powerlaw_2m,banded_10m,uniform_1m_64_local} 2>&1 | grep -E "\|" | tee -a "$LOG"
done
