mkdir -p gpurun_out
for v in "" merge-spmv_b200/variants/lib_evictlast.so; do
for a in "--workload uniform_1m_64" "--workload uniform_1m_64 --cols 8388608" "--workload powerlaw_2m" "--workload banded_10m" "--workload powerlaw_20m --steps 30"; do
MSPMV_LIB=$v timeout 200 python bench.py $a --no-cpu-baseline --no-e2e > gpurun_out/tmp.log 2>&1; python - <<PY
import json
l=open("gpurun_out/tmp.log").read().strip().splitlines()[-1]
try:
    j=json.loads(l); print("${v:-default}", "$a", round(j["ms_per_step"],4),"ms", round(j["value"],1),"GF", round(j["roofline"]["frac"],3))
except Exception as e: print("$v $a FAILED", l[-200:])
PY
done; done
