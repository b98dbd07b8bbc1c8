#!/bin/bash
# compute-sanitizer (memcheck + racecheck on the shared-memory heavy kernels) over small GPU tests.
# The reference never had sanitizer coverage.
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
   python -m pytest tests/test_gpu_corr.py tests/test_gpu_ba.py -q -m gpu -p no:cacheprovider \
   -k "fast_path or fused_multilevel or ba_vs_oracle or structure_only or masked_edges or neighbors_bit_exact or segment_softmax or fused_gru or shared_plan" \
   > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck.log | tail -3
$T 600 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 120 \
   python -m pytest tests/test_gpu_ba.py -q -m gpu -p no:cacheprovider -k "ba_vs_oracle and 4-24" \
   > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | tail -3
