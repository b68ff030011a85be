set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== kbench"; timeout 300 python tools/kbench.py --reps 15 2>&1 | grep -E "step|infer|query|indices|rror"
echo "== kbench slots/node 8"; SHINE_HASH_SLOTS_PER_NODE=8 timeout 300 python tools/kbench.py --quick --reps 15 2>&1 | grep -E "step|infer|rror"
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r02f.err | tee gpurun_out/bench_r02f.json
tail -3 gpurun_out/bench_r02f.err
