set -x
cd $GRAFT_REPO_ROOT
for v in base pf0 pf1 base pf0 pf1; do echo "== variant $v"; SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_$v.so timeout 300 python tools/kbench.py --quick --reps 15 2>&1 | grep -E "step|infer|rror"; done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
bash tools/sanitize.sh > gpurun_out/sanitizer_r02.log 2>&1; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_r02.log
