cd /root/repo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { echo "== $*"; env "$@" MSPMV_BENCH_TRACE=1 $TR --master-port 29530 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-extras $EXTRA 2>&1 | grep -E "trace|ms_per_step" | sed -E 's/.*"ms_per_step": ([0-9.]+).*/ms_per_step \1/' ; }
EXTRA="" run A=1
EXTRA="" run MSPMV_BENCH_SKIP_PARITY=1
EXTRA="--graph off" run A=1
EXTRA="--exchange p2p" run A=1
