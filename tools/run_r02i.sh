set -x
cd $GRAFT_REPO_ROOT
for i in 1 2; do
echo "== random order"; timeout 300 python tools/kbench.py --quick --reps 15 2>&1 | grep -E "step 3xTF32|infer|rror"
echo "== morton-sorted batch"; timeout 300 python tools/kbench.py --quick --reps 15 --sorted 2>&1 | grep -E "step 3xTF32|infer|rror"
done
bash tools/sanitize.sh > gpurun_out/sanitizer_r02.log 2>&1; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_r02.log
