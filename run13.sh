mkdir -p gpurun_out
for v in "" merge-spmv_b200/variants/lib_coldirect.so merge-spmv_b200/variants/lib_coldirect_I11_15.so merge-spmv_b200/variants/lib_I11_15.so merge-spmv_b200/variants/lib_I9_17.so; do
for c in 70 56; do
for a in "--workload uniform_1m_64" "--workload powerlaw_2m" "--workload banded_10m"; do
MSPMV_TILE_CARVEOUT=$c MSPMV_LIB=$v timeout 200 python bench.py $a --no-cpu-baseline --no-e2e --steps 200 2>&1 | tail -1 > gpurun_out/tmp.log; python - <<PY
import json
l=open("gpurun_out/tmp.log").read().strip().splitlines()[-1]
try:
    j=json.loads(l); print("${v:-default}", "carve=$c", "$a", round(j["ms_per_step"],4),"ms", round(j["value"],1),"GF", round(j["roofline"]["frac"],3))
except Exception as e: print("$v $a FAILED", l[-200:])
PY
done; done; done
