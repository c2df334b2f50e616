#!/usr/bin/env bash
# First GPU call of the next round: parity of the opt-in variants, then one timing table of every
# variant on the four single-GPU workloads.  Everything lands in gpurun_out/ (merged back by gpurun).
#
#   gpurun --timeout 1800 -- 'bash tools/round2_sweep.sh'     (~15 min; FULL=1 adds the secondary switches, ~10 min more)
#
# Variants (all off by default; kernel logic already covered on CPU by tests/test_kernel_emu.py):
#   tile_variant=3          thread-blocked gathers, products in registers (csrc/spmv_tile3.cuh)
#   small_fused_tiles=N     single launch for matrices of <= N tiles (config 1 and other small inputs)
#   MSPMV_TILE_PREFETCH=N   L2 prefetch N tiles ahead (measured r01: +2 % banded, -4 % random)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p "$OUT"
PY=python
STEPS=${STEPS:-300}

echo "== parity of the opt-in variants" | tee "$OUT/sweep_r02.txt"
MSPMV_TEST_EXPERIMENTAL=1 timeout 900 $PY -m pytest tests/test_zz_experimental.py -m gpu -q -x 2>&1 | tail -5 | tee -a "$OUT/sweep_r02.txt"

run() {  # label, env assignments..., -- bench args
    local label=$1; shift
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    local line
    line=$(env "${envs[@]}" timeout 600 $PY bench.py --no-cpu-baseline --no-e2e --steps "$STEPS" --warmup 10 "$@" 2>/dev/null | tail -1)
    $PY - "$label" "$line" <<'PYEOF' | tee -a "$OUT/sweep_r02.txt"
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    r = d["roofline"]
    sec = r.get("secondary") or {}
    print(f"{label:34s} {d['config']['workload']:22s} {d['ms_per_step']:.4f} ms  {d['value']:8.1f} GFLOP/s  "
          f"hbm frac {r['frac']:.3f}  gather frac {sec.get('frac', float('nan')):.3f}  "
          f"clk {d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')}")
except Exception as e:
    print(f"{label:34s} FAILED: {e}: {line[:200]}")
PYEOF
}

make -C merge-spmv_b200 -s -j4 variants 2>&1 | tail -3
V=merge-spmv_b200/variants
echo "== timing: ms per CsrMV step (CUDA-graph replay, $STEPS steps; one process per build, both kernels in it)" | tee -a "$OUT/sweep_r02.txt"
lib() {  # label, env assignments... : times tile_variant 2 and 3 of one build on the four workloads
    local label=$1; shift
    env "$@" timeout 900 $PY tools/sweep_lib.py --label "$label" --steps "$STEPS" ${WL:+--workloads "$WL"} > "$OUT/sweep_lib.log" 2>&1
    if grep -qE "ms \|" "$OUT/sweep_lib.log"; then grep -E "^#|\|" "$OUT/sweep_lib.log" | tee -a "$OUT/sweep_r02.txt"
    else echo "$label FAILED:" | tee -a "$OUT/sweep_r02.txt"; tail -8 "$OUT/sweep_lib.log" | tee -a "$OUT/sweep_r02.txt"; fi
}
lib shipped-build
for B in ipt8_12 mbar2us; do lib $B MSPMV_LIB=$V/libmergespmv_$B.so; done
if [ "${FULL:-0}" = 1 ]; then
    for B in ipt7_11 ipt10_14 ipt11_15 mbar20us v3popc v3shflscan v3regs48; do lib $B MSPMV_LIB=$V/libmergespmv_$B.so; done
fi
echo "== shared-memory carve-out (L1 left for gather misses in flight) on the random-column workloads" | tee -a "$OUT/sweep_r02.txt"
# 7 / 11 items per thread are ~14 KB per block: 8-9 blocks fit a 132 KB shared-memory configuration (the gather
# ceiling is 275 G/s there, 252 at 164 KB, 127 at 196 KB: profiles/microbench_r01.txt)
WL=uniform_1m_64,powerlaw_2m
for C in 50 56 62; do
    lib "carve-out $C %" MSPMV_TILE_CARVEOUT=$C
    lib "ipt7_11, carve-out $C %" MSPMV_LIB=$V/libmergespmv_ipt7_11.so MSPMV_TILE_CARVEOUT=$C
done
[ "${FULL:-0}" = 1 ] && lib "carve-out 85 %" MSPMV_TILE_CARVEOUT=85
unset WL
echo "== small matrices (config 1 shape): launch-latency-bound" | tee -a "$OUT/sweep_r02.txt"
run "shipped (3 launches)"        -- --workload cpu_uniform_16k --steps 2000
run "small_fused_tiles=4096"      -- --workload cpu_uniform_16k --steps 2000 --option small_fused_tiles=4096
run "shipped, eager launches"     -- --workload cpu_uniform_16k --steps 2000 --graph off
run "fused, eager launches"       -- --workload cpu_uniform_16k --steps 2000 --graph off --option small_fused_tiles=4096

echo "== same-box comparators through the C++ driver (merge CsrMV | cusparseSpMV ALG2 | toolkit cub::DeviceSpmv)" | tee -a "$OUT/sweep_r02.txt"
D=merge-spmv_b200/gpu_spmv
( timeout 600 $D --uniform=64 --rows=1048576 --values=random --randx --cusparse --cub
  timeout 600 $D --powerlaw=1000000 --rows=2000000 --nnz=200000000 --fp32 --values=random --randx --cusparse --cub
  timeout 600 $D --banded=3 --rows=10000000 --values=random --randx --cusparse --cub ) 2>&1 \
    | grep -E "CsrMV|SpMV|PASS|FAIL|avg ms" | tee -a "$OUT/sweep_r02.txt" | tail -40

# multi-GPU (run this script under `gpurun --gpus 2`): NCCL all_gather + fold vs the NVLink peer-memory
# exchange kernel, and the solver-style step with the y all_gather
NGPU=$(nvidia-smi -L 2>/dev/null | wc -l)
if [ "$NGPU" -ge 2 ]; then
    echo "== $NGPU GPUs: carry exchange" | tee -a "$OUT/sweep_r02.txt"
    MSPMV_TEST_EXPERIMENTAL=1 timeout 900 $PY -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -3 | tee -a "$OUT/sweep_r02.txt"
    for X in nccl p2p; do
        line=$(timeout 900 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node "$NGPU" --master-addr 127.0.0.1 \
               --master-port 29517 bench.py --gpus "$NGPU" --steps "$STEPS" --warmup 10 --no-e2e --exchange $X 2>/dev/null | tail -1)
        echo "exchange=$X: $line" | cut -c1-400 | tee -a "$OUT/sweep_r02.txt"
    done
    line=$(timeout 900 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node "$NGPU" --master-addr 127.0.0.1 \
           --master-port 29518 bench.py --gpus "$NGPU" --steps "$STEPS" --warmup 10 --no-e2e --gather-y 2>/dev/null | tail -1)
    echo "gather-y: $line" | cut -c1-400 | tee -a "$OUT/sweep_r02.txt"
    echo "-- C++ driver, one process driving $NGPU devices (peer-memory carry exchange, no NCCL)" | tee -a "$OUT/sweep_r02.txt"
    timeout 900 merge-spmv_b200/gpu_spmv --uniform=64 --rows=$((1048576 * NGPU)) --cols=1048576 --values=random --randx --gpus=$NGPU 2>&1 \
        | grep -E "CsrMV|PASS|FAIL|avg ms" | tee -a "$OUT/sweep_r02.txt"
    for P in "" "--e2e-pipeline"; do
        line=$(timeout 900 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node "$NGPU" --master-addr 127.0.0.1 \
               --master-port 29519 bench.py --gpus "$NGPU" --steps 100 --warmup 10 $P 2>/dev/null | tail -1)
        echo "e2e ${P:-sequential}: $($PY -c "import json,sys; d=json.loads(sys.argv[1]); print(d['e2e'])" "$line" 2>/dev/null)" | tee -a "$OUT/sweep_r02.txt"
    done
fi

# one ncu capture of the candidate kernel next to the shipped one (banded: the issue-bound case)
if command -v ncu >/dev/null; then
    for V in 2 3; do
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_tile -c 1 \
            -o "$OUT/ncu_r02_banded_v$V" -f $PY bench.py --workload banded_10m --steps 2 --warmup 3 \
            --no-cpu-baseline --no-e2e --graph off --option tile_variant=$V > "$OUT/ncu_r02_banded_v$V.log" 2>&1
    done
fi
echo "done" | tee -a "$OUT/sweep_r02.txt"
