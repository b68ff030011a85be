# compute-sanitizer over one small fused step + query fwd/bwd + adam (memcheck, racecheck, synccheck)
set -x
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged_batch_sizes and 17 or fused_step_matches_oracle and 2-True or adam_kernel" 2>&1 | tail -6
done
