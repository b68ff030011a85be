set -x
cd $GRAFT_REPO_ROOT
echo "== perm+dynamic sorted"; timeout 300 python tools/kbench.py --quick --reps 15 --sorted 2>&1 | grep -E "grouped|infer|rror"
echo "== perm+static sorted"; SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_permstatic.so timeout 300 python tools/kbench.py --quick --reps 15 --sorted 2>&1 | grep -E "grouped|infer|rror"
echo "== perm+dynamic random"; timeout 300 python tools/kbench.py --quick --reps 15 2>&1 | grep -E "step 3xTF32|infer|rror"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
