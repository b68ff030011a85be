set -x
cd $GRAFT_REPO_ROOT
for v in cur cpa1 pf2 cur cpa1 pf2; do echo "== $v sorted"; SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_$v.so timeout 300 python tools/kbench.py --quick --reps 15 --sorted 2>&1 | grep -E "grouped|rror"; done
timeout 900 python bench.py --steps 20 --warmup 5 --no-hbm-leg 2>gpurun_out/bench_r02t.err | tee gpurun_out/bench_r02t.json | cut -c1-300
tail -3 gpurun_out/bench_r02t.err
