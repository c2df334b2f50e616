mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for w in uniform_1m_64 powerlaw_2m banded_10m; do
timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 200 > gpurun_out/bench_$w.log 2>&1; python - <<PY
import json
l=open("gpurun_out/bench_$w.log").read().strip().splitlines()[-1]
try:
    j=json.loads(l); print("$w", round(j["ms_per_step"],4),"ms", round(j["value"],1),"GF", round(j["roofline"]["frac"],3))
except Exception as e: print("$w", l[-300:])
PY
done
