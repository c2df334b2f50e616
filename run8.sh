mkdir -p gpurun_out
./merge-spmv_b200/bin/microbench > gpurun_out/microbench_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
for w in uniform_1m_64 banded_10m powerlaw_2m; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_tile -s 4 -c 1 -o gpurun_out/prof9_$w -f python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$w.log 2>&1; echo "ncu $w rc=$?"
done
