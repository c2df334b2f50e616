#!/usr/bin/env bash
# Multi-GPU validation + timing on N GPUs of one box:  gpurun --gpus N --timeout 1500 -- 'bash tools/mg_sweep.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
N=$(nvidia-smi -L 2>/dev/null | wc -l)
LOG="$OUT/mg_sweep_n$N.txt"; : > "$LOG"
PY=python
TR="$PY -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== $N GPUs: NCCL and NVLink carry exchange parity tests, eager and graph-captured" | tee -a "$LOG"
timeout 900 $PY -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -4 | tee -a "$LOG"
echo "== C-ABI multi-GPU session + C++ driver tests on $N devices" | tee -a "$LOG"
timeout 900 $PY -m pytest tests -m gpu -q -x -k "mg_session or gpu_driver_self_check" 2>&1 | tail -4 | tee -a "$LOG"
summ() { $PY -c "
import json,sys
d=json.loads(sys.argv[1])
print(f\"{sys.argv[2]:28s} {d['config']['workload']:16s} {d['ms_per_step']:.4f} ms  {d['value']:9.1f} GFLOP/s  frac {d['roofline']['frac']:.3f}  parity {d['parity']['ok']} {d['parity']['max_rel']:.2e}  e2e {(d['e2e'] or {}).get('value')}  clk {d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')}\")
for e in d.get('extra_workloads') or []:
    print(f\"{'':28s} {e['workload']:16s} {e['ms_per_step']:.4f} ms  {e['value']:9.1f} GFLOP/s  frac {e['roofline_frac']:.3f}  parity {e['parity']['ok']} {e['parity']['max_rel']:.2e}\")
" "$1" "$2" 2>&1 | tee -a "$LOG"; }
echo "== bench.py --gpus $N (weak: one config-2 shard per GPU)" | tee -a "$LOG"
for X in nccl p2p; do
    line=$(timeout 900 $TR --master-port 29517 bench.py --gpus $N --steps 300 --warmup 10 --no-e2e --no-extras --exchange $X 2>"$OUT/mg_err_$X.txt" | tail -1)
    summ "$line" "exchange=$X" || tail -5 "$OUT/mg_err_$X.txt" | tee -a "$LOG"
done
line=$(timeout 900 $TR --master-port 29518 bench.py --gpus $N --steps 300 --warmup 10 --no-e2e --no-extras --gather-y 2>/dev/null | tail -1)
summ "$line" "nccl + gather-y"
echo "== the driver's own command (default options: e2e pipelined, extras = config 5 strong)" | tee -a "$LOG"
line=$(timeout 1200 $TR --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 2>"$OUT/mg_err_default.txt" | tail -1)
echo "$line" > "$OUT/bench_r02_n$N.json"
summ "$line" "default" || tail -5 "$OUT/mg_err_default.txt" | tee -a "$LOG"
line=$(timeout 900 $TR --master-port 29520 bench.py --gpus $N --steps 20 --warmup 3 --no-extras --e2e-sequential 2>/dev/null | tail -1)
summ "$line" "e2e sequential"
echo "== reference arm under torchrun (rank 0 only)" | tee -a "$LOG"
timeout 900 $TR --master-port 29521 bench.py --impl reference --gpus $N --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-600 | tee -a "$LOG"
echo "== C++ driver, one process driving $N devices (peer-memory carry exchange, no NCCL, no Python)" | tee -a "$LOG"
timeout 900 merge-spmv_b200/gpu_spmv --uniform=64 --rows=$((1048576 * N)) --cols=1048576 --values=random --randx --gpus=$N 2>&1 \
    | grep -E "CsrMV|PASS|FAIL|avg ms|rror|host buffers" | tee -a "$LOG"
echo done | tee -a "$LOG"
