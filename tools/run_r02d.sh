set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05_train" 2>&1 | tail -25
echo "== kbench"; timeout 300 python tools/kbench.py --quick --reps 15 2>&1 | grep -E "step|infer|rror"
