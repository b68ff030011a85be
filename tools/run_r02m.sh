set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --steps 20 --warmup 5 --no-hbm-leg 2>gpurun_out/bench_r02m.err | tee gpurun_out/bench_r02m.json | cut -c1-400
tail -3 gpurun_out/bench_r02m.err
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sdf_fused_kernel -c 1 -o gpurun_out/r02_step_c2_grouped python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hbm-leg --cuda-profiler > gpurun_out/ncu_c2g.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
