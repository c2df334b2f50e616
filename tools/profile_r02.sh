#!/usr/bin/env bash
# Round-2 profiling evidence: the launch list of the default bench command, and one `ncu --set full`
# capture of the pipe kernel per headline workload.  gpurun --timeout 1200 -- 'bash tools/profile_r02.sh'
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmv_pipe|tile_search|carry|diagonal" -c 12 --csv --log-file "$OUT/launches_r02.csv" \
    python bench.py --steps 5 --warmup 3 --graph off --no-cpu-baseline --no-e2e --no-extras > "$OUT/launches_r02.log" 2>&1
for W in ${WORKLOADS:-uniform_1m_64 banded_10m powerlaw_2m}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_pipe -c 1 \
        -o "$OUT/ncu_r02_$W" -f python bench.py --workload $W --steps 2 --warmup 3 \
        --no-cpu-baseline --no-e2e --no-extras --graph off > "$OUT/ncu_r02_$W.log" 2>&1
    tail -1 "$OUT/ncu_r02_$W.log" | cut -c1-200
done
