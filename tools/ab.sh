for r in 0 1 0 1; do echo "== fused replicas $r"; SHINE_FUSED_REPLICAS=$r timeout 200 python tools/kbench.py 2>/dev/null | grep -E "step 3xTF32"; done
for t in 64 1024; do echo "== replicas target $t"; SHINE_REPLICA_TARGET=$t timeout 200 python tools/kbench.py 2>/dev/null | grep -E "step 3xTF32"; done
