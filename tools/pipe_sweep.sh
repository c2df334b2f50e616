#!/usr/bin/env bash
# Round-2 timing sweep of the pipe kernel (csrc/spmv_pipe.cuh): launch modes, compile-time shapes, shared-memory budgets.
#   gpurun --timeout 1500 -- 'bash tools/pipe_sweep.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p "$OUT"
PY=python
STEPS=${STEPS:-200}
LOG="$OUT/pipe_sweep.txt"
: > "$LOG"
V=merge-spmv_b200/variants
lib() {  # label, options, env assignments...
    local label=$1; shift
    local opts=$1; shift
    env "$@" timeout 900 $PY tools/sweep_lib.py --label "$label" --steps "$STEPS" --options "$opts" ${WL:+--workloads "$WL"} > "$OUT/sweep_lib.log" 2>&1
    if grep -qE "ms \|" "$OUT/sweep_lib.log"; then grep -E "^#|\|" "$OUT/sweep_lib.log" | tee -a "$LOG"
    else echo "$label FAILED:" | tee -a "$LOG"; tail -12 "$OUT/sweep_lib.log" | tee -a "$LOG"; fi
}
if [ "${SKIP_TESTS:-0}" != 1 ]; then
    echo "== gpu tests" | tee -a "$LOG"
    timeout 1200 $PY -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee -a "$LOG"
fi
echo "== shipped build: one launch vs search kernel + pipe kernel; both kernel shapes forced" | tee -a "$LOG"
lib shipped "engine=pipe;engine=pipe,pipe_search=0;engine=pipe,pipe_config=1;engine=pipe,pipe_config=2"
echo "== pipe: compile-time variants" | tee -a "$LOG"
for B in ${VARIANTS:-p_base p_c1 p_ah p_ah_c1 p_ipt11 p_ipt11_c1 p_ah_ipt7_c1 p_nw8_c1 p_nw2 p_nw2_c1 p_nw2_ah_c1 p_nw1 p_nw1_ipt15 p_nw2_ipt13}; do
    [ -f $V/libmergespmv_$B.so ] && lib $B "engine=pipe" MSPMV_LIB=$V/libmergespmv_$B.so
done
echo "== pipe: shared-memory budget per SM (rest is L1) / resident blocks" | tee -a "$LOG"
for KB in ${SMEM_KBS:-96 128 192}; do lib "smem ${KB}KB" "engine=pipe" MSPMV_PIPE_SMEM_KB=$KB; done
echo "== small matrix (config 1 shape)" | tee -a "$LOG"
WL=cpu_uniform_16k STEPS=2000 lib "small" "engine=pipe;engine=pipe,pipe_search=0"
echo done | tee -a "$LOG"
