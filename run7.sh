mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_reference.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_reference.log
for w in powerlaw_2m banded_10m; do timeout 300 python bench.py --workload $w --no-e2e > gpurun_out/bench_$w.log 2>&1; tail -1 gpurun_out/bench_$w.log; done
nproc; lscpu | grep -E "Model name|Socket|Thread|Core" 
