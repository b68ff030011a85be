for spn in 2 4 8; do echo "== slots/node $spn"; SHINE_HASH_SLOTS_PER_NODE=$spn timeout 200 python tools/kbench.py 2>/dev/null | grep -E "step 3xTF32|infer 3x|query_fwd|get_indices"; done
