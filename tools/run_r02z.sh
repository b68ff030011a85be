set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== sorted"; timeout 300 python tools/kbench.py --quick --reps 15 --sorted 2>&1 | grep -E "step 3xTF32|infer|rror"
echo "== random"; timeout 300 python tools/kbench.py --quick --reps 15 2>&1 | grep -E "step 3xTF32|infer|rror"
timeout 600 python bench.py --steps 20 --warmup 5 --no-hbm-leg --no-cpu-baseline 2>gpurun_out/bench_r02z.err | tee gpurun_out/bench_r02z.json | cut -c1-200
tail -2 gpurun_out/bench_r02z.err
