set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/bench_r02_n8_graph.err | grep "^{" | tee gpurun_out/bench_r02_n8_graph.json | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 2>gpurun_out/bench_r02_n4_graph.err | grep "^{" | tee gpurun_out/bench_r02_n4_graph.json | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --global-points 1048576 2>gpurun_out/bench_r02_c5_graph.err | grep "^{" | tee gpurun_out/bench_r02_c5_graph.json | cut -c1-250
