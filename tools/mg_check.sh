#!/usr/bin/env bash
# the driver's own N-GPU command (20 steps, 3 warm-up) + a 300-step run:  gpurun --gpus N -- 'bash tools/mg_check.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
N=$(nvidia-smi -L 2>/dev/null | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
summ() { python -c "
import json,sys
d=json.loads(sys.argv[1])
print(f\"{sys.argv[2]:22s} {d['config']['workload']:16s} {d['ms_per_step']:.4f} ms {d['value']:9.1f} GFLOP/s frac {d['roofline']['frac']:.3f} parity {d['parity']['ok']} e2e {(d['e2e'] or {}).get('value')} clk {d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')}\")
for e in d.get('extra_workloads') or []:
    print(f\"{'':22s} {e['workload']:16s} {e['ms_per_step']:.4f} ms {e['value']:9.1f} GFLOP/s frac {e['roofline_frac']:.3f} parity {e['parity']['ok']} {e['parity']['max_rel']:.2e}\")
" "$1" "$2"; }
line=$(MSPMV_BENCH_TRACE=1 timeout 1200 $TR --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 2>"$OUT/mg_check_err.txt" | tail -1)
echo "$line" > "$OUT/bench_r02_n$N.json"
summ "$line" "default 20 steps" || tail -20 "$OUT/mg_check_err.txt"
grep trace "$OUT/mg_check_err.txt" | cut -c1-260 | head -6
line=$(timeout 900 $TR --master-port 29520 bench.py --gpus $N --steps 300 --warmup 10 --no-extras ${EXTRA:-} 2>/dev/null | tail -1)
summ "$line" "300 steps"
