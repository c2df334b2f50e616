mkdir -p gpurun_out
for w in uniform_1m_64 banded_10m powerlaw_2m; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_tile -s 4 -c 1 -o gpurun_out/prof7_$w -f python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$w.log 2>&1; echo "ncu $w rc=$?"
done
