set -x
cd $GRAFT_REPO_ROOT
bash tools/sanitize.sh > gpurun_out/sanitizer_r02.log 2>&1; tail -20 gpurun_out/sanitizer_r02.log
timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -q -s -k "cuda_update" 2>&1 | grep -E "launches|passed|failed"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "beats_class_surface" 2>&1 | grep -E "eikonal step|passed|failed"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-leg --cuda-profiler > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sdf_fused_kernel -c 1 -o gpurun_out/r02_step_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hbm-leg --cuda-profiler > gpurun_out/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdf_fused_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_step_hbm python bench.py --hbm-only --steps 5 > gpurun_out/ncu_hbm.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r02g.err | tee gpurun_out/bench_r02g.json | cut -c1-200
ls gpurun_out | tail -5
