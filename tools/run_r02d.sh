set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/debug_tc.py 2>&1 | grep -v Warn | tail -40
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05_train" 2>&1 | tail -8
echo "== kbench"; timeout 300 python tools/kbench.py --quick --reps 15 2>&1 | grep -E "step|infer|rror"
