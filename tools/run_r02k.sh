set -x
cd $GRAFT_REPO_ROOT
for v in cur nored; do
echo "== $v random"; SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_$v.so timeout 300 python tools/kbench.py --quick --reps 15 2>&1 | grep -E "step 3xTF32|rror"
echo "== $v sorted"; SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_$v.so timeout 300 python tools/kbench.py --quick --reps 15 --sorted 2>&1 | grep -E "step 3xTF32|rror"
done
