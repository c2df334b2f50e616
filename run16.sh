mkdir -p gpurun_out
for v in merge-spmv_b200/variants/lib_adaptive.so; do
for a in "--workload uniform_1m_64" "--workload powerlaw_2m" "--workload banded_10m" "--workload uniform_1m_64_local" "--workload powerlaw_20m --steps 20"; do
MSPMV_LIB=$v timeout 200 python bench.py $a --no-cpu-baseline --no-e2e --steps 300 2>&1 | tail -1 > gpurun_out/tmp.log; python - <<PY
import json
l=open("gpurun_out/tmp.log").read().strip().splitlines()[-1]
try:
    j=json.loads(l); print("${v:-default}", "$a", round(j["ms_per_step"],4),"ms", round(j["value"],1),"GF", round(j["roofline"]["frac"],3))
except Exception as e: print("$v $a FAILED", l[-200:])
PY
done; done
MSPMV_LIB=merge-spmv_b200/variants/lib_adaptive.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
