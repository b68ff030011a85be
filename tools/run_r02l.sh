set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped" 2>&1 | tail -15
for i in 1 2; do
echo "== random order"; timeout 300 python tools/kbench.py --quick --reps 15 2>&1 | grep -E "step 3xTF32|rror"
echo "== morton-sorted batch"; timeout 300 python tools/kbench.py --quick --reps 15 --sorted 2>&1 | grep -E "step 3xTF32|infer|rror"
done
