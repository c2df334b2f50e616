mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 ./merge-spmv_b200/bin/microbench 2>&1 | grep -E "table=(4096|16384) " > gpurun_out/microbench_small.log; cat gpurun_out/microbench_small.log | grep uniform
for w in uniform_1m_64 banded_10m; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_stream -s 4 -c 1 -o gpurun_out/prof_$w -f python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$w.log 2>&1; echo "ncu $w rc=$?"
done
