set -x
cd $GRAFT_REPO_ROOT
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/exchange_lat.py 2>&1 | grep -E "world|rror" | tee gpurun_out/exchange_lat_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_r02_n${N}_auto.err | grep "^{" | tee gpurun_out/bench_r02_n${N}_auto.json | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --global-points 1048576 2>gpurun_out/bench_r02_c5.err | grep "^{" | tee gpurun_out/bench_r02_c5.json | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 2>gpurun_out/bench_r02_n4_auto.err | grep "^{" | tee gpurun_out/bench_r02_n4_auto.json | cut -c1-330
