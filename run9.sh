mkdir -p gpurun_out
for g in on off; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 300 --warmup 10 --graph $g --no-e2e > gpurun_out/bench_2gpu_$g.log 2>&1; echo "bench2 graph=$g rc=$?"; tail -1 gpurun_out/bench_2gpu_$g.log | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['value'], j['gpu_launches'])"
done
timeout 300 python bench.py --graph on --no-e2e --no-cpu-baseline --steps 300 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('1gpu graph', j['ms_per_step'], j['value'], j['gpu_launches'])"
timeout 300 python bench.py --graph off --no-e2e --no-cpu-baseline --steps 300 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('1gpu eager', j['ms_per_step'], j['value'], j['gpu_launches'])"
