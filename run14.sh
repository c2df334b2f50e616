mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "known_answers or golden_vectors or misaligned or temp_storage" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/sanitizer.log; grep -c "Invalid" gpurun_out/sanitizer.log
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log
