#!/usr/bin/env bash
# End-of-round validation on one B200:  gpurun --timeout 1800 -- 'bash tools/final_check.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"; LOG="$OUT/final_check.txt"; : > "$LOG"
echo "== pytest -m gpu" | tee -a "$LOG"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee -a "$LOG"
echo "== smoke()" | tee -a "$LOG"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a "$LOG"
echo "== bench.py (default command)" | tee -a "$LOG"
timeout 900 python bench.py > "$OUT/bench_r02_n1.json" 2> "$OUT/bench_r02_n1.err"
python - "$OUT/bench_r02_n1.json" <<'PY' 2>&1 | tee -a "$LOG"
import json, sys
d = json.load(open(sys.argv[1]))
print(f"{d['config']['workload']:16s} {d['ms_per_step']:.4f} ms {d['value']:8.1f} GFLOP/s  hbm frac {d['roofline']['frac']:.3f}  port frac {d['roofline']['secondary']['frac']:.3f}  "
      f"parity {d['parity']['ok']} {d['parity']['max_rel']:.2e}  e2e {d['e2e']['value']:.1f}  cpu {d['cpu_baseline']['value']:.2f} x{d['cpu_baseline']['cores']}  launches {d['gpu_launches']}  clk {d['clocks']}")
for e in d["extra_workloads"]:
    print(f"{e['workload']:16s} {e['ms_per_step']:.4f} ms {e['value']:8.1f} GFLOP/s  hbm frac {e['roofline_frac']:.3f}  parity {e['parity']['ok']} {e['parity']['max_rel']:.2e}")
PY
echo "== bench.py --impl reference --steps 5 --warmup 3" | tee -a "$LOG"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-400 | tee -a "$LOG"
echo "== bench.py --steps 20 --warmup 3 (the driver's step count)" | tee -a "$LOG"
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])" | tee -a "$LOG"
echo "== small matrix (config 1 shape), 2000 steps" | tee -a "$LOG"
timeout 300 python bench.py --workload cpu_uniform_16k --steps 2000 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])" | tee -a "$LOG"
echo "== ncu launch list of the CsrMV steps (eager launches)" | tee -a "$LOG"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmv_pipe|tile_search|carry|diagonal" -c 12 --csv \
    --log-file "$OUT/launches_r02.csv" python bench.py --steps 5 --warmup 3 --graph off --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1
grep -c spmv_pipe "$OUT/launches_r02.csv" | tee -a "$LOG"
echo "== drivers: merge CsrMV | cusparseSpMV ALG2 | toolkit cub::DeviceSpmv" | tee -a "$LOG"
D=merge-spmv_b200/gpu_spmv
timeout 600 $D --uniform=64 --rows=1048576 --values=random --randx --cusparse --cub > "$OUT/driver_r02_uniform.txt" 2>&1
timeout 600 $D --powerlaw=1000000 --rows=2000000 --nnz=200000000 --fp32 --values=random --randx --cusparse --cub > "$OUT/driver_r02_powerlaw.txt" 2>&1
timeout 600 $D --banded=3 --rows=10000000 --values=random --randx --cusparse --cub > "$OUT/driver_r02_banded.txt" 2>&1
timeout 300 $D --uniform=32 --rows=16384 --values=random --randx --cusparse --cub > "$OUT/driver_r02_config1.txt" 2>&1
grep -hE "CsrMV|SpMV|PASS|FAIL|avg ms" "$OUT"/driver_r02_*.txt | grep -v Invoking | tee -a "$LOG"
if [ "${SANITIZE:-1}" = 1 ]; then
    echo "== compute-sanitizer memcheck on the edge-case parity tests" | tee -a "$LOG"
    timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
        -k "known_answers or misaligned or temp_storage or alpha_beta" > "$OUT/memcheck_r02.txt" 2>&1
    echo "exit $?" | tee -a "$LOG"; grep -E "ERROR SUMMARY|passed|failed" "$OUT/memcheck_r02.txt" | tail -3 | tee -a "$LOG"
fi
echo done | tee -a "$LOG"
