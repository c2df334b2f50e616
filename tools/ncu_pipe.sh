#!/usr/bin/env bash
# ncu --set full captures of the pipe kernel on the listed workloads (default: banded + uniform)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p "$OUT"
TAG=${TAG:-r02}
for W in ${WORKLOADS:-banded_10m uniform_1m_64}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_pipe -c 1 \
        -o "$OUT/ncu_${TAG}_$W" -f python bench.py --workload $W --steps 2 --warmup 3 \
        --no-cpu-baseline --no-e2e --graph off > "$OUT/ncu_${TAG}_$W.log" 2>&1
    tail -2 "$OUT/ncu_${TAG}_$W.log"
done
