set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for v in 000 100 010 001 110 101 011 111; do echo "== variant $v (sector,cpasync,dw3tmem)"; SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_$v.so timeout 200 python tools/kbench.py --quick --reps 15 2>/dev/null | grep -E "step|infer"; done
for v in 000 111; do
SHINE_B200_LIB=$PWD/tools/variants/libshine_b200_$v.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:sdf_fused_kernel -c 1 -o gpurun_out/r02_train_$v python tools/kbench.py --quick --reps 3 > gpurun_out/ncu_$v.log 2>&1
done
ls -la gpurun_out
