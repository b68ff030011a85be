set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
bash tools/sanitize.sh > gpurun_out/sanitizer_r02.log 2>&1; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_r02.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python bench.py --eager --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-leg --cuda-profiler > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sdf_fused_kernel -c 1 -o gpurun_out/r02_step_c2_grouped python bench.py --eager --steps 2 --warmup 3 --no-cpu-baseline --no-hbm-leg --cuda-profiler > gpurun_out/ncu_c2g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sdf_fused_kernel -c 1 -o gpurun_out/r02_step_c2 python bench.py --eager --batch-order random --steps 2 --warmup 3 --no-cpu-baseline --no-hbm-leg --cuda-profiler > gpurun_out/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdf_fused_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_step_hbm python bench.py --hbm-only --steps 5 > gpurun_out/ncu_hbm.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r02x.err | tee gpurun_out/bench_r02x.json | cut -c1-200
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_r02x_ref.err | tee gpurun_out/bench_r02x_ref.json | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ls gpurun_out | tail -5
