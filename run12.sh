mkdir -p gpurun_out
echo skip-pytest
for a in "--workload uniform_1m_64" "--workload powerlaw_2m" "--workload banded_10m"; do
timeout 200 python bench.py $a --no-cpu-baseline --no-e2e --steps 300 2>&1 | tail -1 > gpurun_out/tmp.log; python - <<PY
import json
l=open("gpurun_out/tmp.log").read().strip().splitlines()[-1]
try:
    j=json.loads(l); print("$a", round(j["ms_per_step"],4),"ms", round(j["value"],1),"GF", round(j["roofline"]["frac"],3), j["gpu_launches"])
except Exception as e: print("$a FAILED", l[-200:])
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep mspmv gpurun_out/launches_r01b.csv | python -c "
import sys,csv,collections
agg=collections.defaultdict(list)
for r in csv.reader(sys.stdin):
    agg[r[4].split('(')[0][:50]].append(float(r[-1].replace(',','')))
for k,v in agg.items(): print(k, len(v), 'avg us', round(sum(v)/len(v)/1e3,2))
"
