set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_partition.py tests/test_gpu_loop.py -m gpu -x -q 2>&1 | tail -4
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_r02_n${N}_graph.err | grep "^{" | tee gpurun_out/bench_r02_n${N}_graph.json | cut -c1-250
tail -3 gpurun_out/bench_r02_n2_graph.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-hbm-leg --no-cpu-baseline 2>gpurun_out/bench_r02v.err | tee gpurun_out/bench_r02v.json | cut -c1-250
tail -3 gpurun_out/bench_r02v.err
