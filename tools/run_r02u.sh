set -x
cd $GRAFT_REPO_ROOT
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_r02_n${N}_rot.err | grep "^{" | tee gpurun_out/bench_r02_n${N}_rot.json | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench_r02_n2_rot.err | grep "^{" | tee gpurun_out/bench_r02_n2_rot.json | cut -c1-330
