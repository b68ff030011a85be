set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for v in cur cpa1 cur cpa1; do echo "== $v sorted"; SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_$v.so timeout 300 python tools/kbench.py --quick --reps 15 --sorted 2>&1 | grep -E "grouped|rror"; done
echo "== cur random"; SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_cur.so timeout 300 python tools/kbench.py --quick --reps 15 2>&1 | grep -E "step 3xTF32|rror"
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/exchange_lat.py 2>&1 | grep -E "world|rror" | tee gpurun_out/exchange_lat_n$N.txt
for ex in auto nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --exchange $ex 2>gpurun_out/bench_r02_n${N}_$ex.err | grep "^{" | tee gpurun_out/bench_r02_n${N}_$ex.json | cut -c1-200
done
