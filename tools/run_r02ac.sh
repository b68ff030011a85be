set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r02ac.err | tee gpurun_out/bench_r02ac.json | cut -c1-200
bash tools/sanitize.sh > gpurun_out/sanitizer_r02.log 2>&1; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_r02.log
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sdf_fused_kernel -c 1 -o gpurun_out/r02_step_c2_grouped python bench.py --eager --steps 2 --warmup 3 --no-cpu-baseline --no-hbm-leg --cuda-profiler > gpurun_out/ncu_c2g.log 2>&1
