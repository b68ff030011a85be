set -x
cd $GRAFT_REPO_ROOT
N=2
timeout 900 python -m pytest tests/test_gpu_partition.py tests/test_gpu_loop.py -m gpu -x -q -k "two_gpu or abi_launches" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/exchange_lat.py 2>&1 | grep -E "world|rror" | tee gpurun_out/exchange_lat_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_r02_n${N}_auto.err | grep "^{" | tee gpurun_out/bench_r02_n${N}_auto.json | cut -c1-200
