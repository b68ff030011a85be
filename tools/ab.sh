echo "== big map"; timeout 200 python tools/kbench.py --frames 60 --step-m 3 --leaf-vox 0.05 --points 1048576 2>/dev/null | grep -E "table MB|step 3xTF32|infer 3x|query_|adam|zero"
for t in 64 128 512 2048; do echo "== replica target $t"; SHINE_REPLICA_TARGET=$t timeout 200 python tools/kbench.py 2>/dev/null | grep -E "step 3xTF32"; done
