set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_partition.py -m gpu -x -q -k "two_gpu or abi_launches" 2>&1 | tail -5
for ex in p2p nccl p2p; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --exchange $ex 2>gpurun_out/bench_r02_n2_$ex.err | grep "^{" | tee gpurun_out/bench_r02_n2_$ex.json | cut -c1-330
done
