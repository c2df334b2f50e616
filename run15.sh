mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1
for w in uniform_1m_64 banded_10m powerlaw_2m; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_tile -s 4 -c 1 -o gpurun_out/prof_final_$w -f python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$w.log 2>&1; echo "ncu $w rc=$?"
done
timeout 300 ./merge-spmv_b200/gpu_spmv --uniform=64 --rows=1048576 --values=random --randx --cusparse > gpurun_out/driver_uniform.log 2>&1
timeout 300 ./merge-spmv_b200/gpu_spmv --banded=3 --rows=10000000 --values=random --randx --cusparse > gpurun_out/driver_banded.log 2>&1
timeout 300 ./merge-spmv_b200/gpu_spmv --powerlaw=1000000 --rows=2000000 --nnz=200000000 --fp32 --values=random --randx --cusparse > gpurun_out/driver_powerlaw.log 2>&1
grep -h -A3 -E "^Merge-based|^cuSPARSE" gpurun_out/driver_*.log | grep -E "fp|PASS|FAIL|Merge|cuSPARSE"
