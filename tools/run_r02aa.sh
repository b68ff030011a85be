set -x
cd $GRAFT_REPO_ROOT
for v in static dyn; do for r in 0 1; do
echo "== $v replicas=$r sorted"; SHINE_GROUPED_REPLICAS=$r SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_$v.so timeout 300 python tools/kbench.py --quick --reps 15 --sorted 2>&1 | grep -E "grouped|infer|rror"
done; done
