set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do echo "== sorted"; timeout 300 python tools/kbench.py --quick --reps 15 --sorted 2>&1 | grep -E "grouped|infer|rror"; done
