set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv
nvidia-smi topo -m 2>&1 | head -12
timeout 900 python -m pytest tests/test_gpu_loop.py tests/test_gpu_partition.py -m gpu -x -q -k "two_gpu" 2>&1 | tail -15
for ex in nccl p2p; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex 2>gpurun_out/bench_r02_n2_$ex.err | tee gpurun_out/bench_r02_n2_$ex.json
tail -3 gpurun_out/bench_r02_n2_$ex.err
done
echo "== host step variants (1 GPU)"; timeout 300 python tools/zero_copy.py 2>&1 | tail -5
