set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== kbench default (g2)"; timeout 300 python tools/kbench.py 2>&1 | grep -vi warn | tail -16
echo "== kbench g4"; SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_g4.so timeout 300 python tools/kbench.py 2>&1 | grep -E "step 3xTF32|infer 3x|query_"
echo "== big map"; timeout 300 python tools/kbench.py --frames 60 --step-m 3 --leaf-vox 0.05 --points 1048576 2>/dev/null | grep -E "table MB|step 3xTF32|infer 3x|query_|adam|zero"
echo "== zero copy"; timeout 300 python tools/zero_copy.py 2>&1 | tail -6
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r02a.err | tee gpurun_out/bench_r02a.json
tail -5 gpurun_out/bench_r02a.err
