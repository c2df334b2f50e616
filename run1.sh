mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 ./merge-spmv_b200/bin/microbench > gpurun_out/microbench.log 2>&1; echo "microbench rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.log
timeout 300 python bench.py --engine tile --no-cpu-baseline --no-e2e > gpurun_out/bench_tile.log 2>&1; tail -1 gpurun_out/bench_tile.log
timeout 300 python bench.py --workload powerlaw_2m --no-cpu-baseline --no-e2e > gpurun_out/bench_powerlaw.log 2>&1; tail -1 gpurun_out/bench_powerlaw.log
timeout 300 python bench.py --workload banded_10m --no-cpu-baseline --no-e2e > gpurun_out/bench_banded.log 2>&1; tail -1 gpurun_out/bench_banded.log
