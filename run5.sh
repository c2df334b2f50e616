mkdir -p gpurun_out
for v in "" merge-spmv_b200/variants/lib_T128_I11.so merge-spmv_b200/variants/lib_T128_I13.so merge-spmv_b200/variants/lib_T128_I15.so merge-spmv_b200/variants/lib_T256_I9.so; do
for c in 70 62; do
for w in uniform_1m_64 powerlaw_2m banded_10m; do
MSPMV_TILE_CARVEOUT=$c MSPMV_LIB=$v timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 200 > gpurun_out/tmp.log 2>&1; python - <<PY
import json
l=open("gpurun_out/tmp.log").read().strip().splitlines()[-1]
try:
    j=json.loads(l); print("${v:-T128_I9}", "carve=$c", "$w", round(j["ms_per_step"],4),"ms", round(j["value"],1),"GF", round(j["roofline"]["frac"],3))
except Exception as e: print("$v $w FAILED", l[-200:])
PY
done; done; done
