#!/bin/bash
# compute-sanitizer (memcheck + racecheck on the shared-memory heavy kernels) over small GPU tests.
# The reference never had sanitizer coverage.
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
   python -m pytest tests/test_gpu_corr.py tests/test_gpu_ba.py -q -m gpu -p no:cacheprovider \
   -k "fast_path or split_precision or fused_multilevel or ba_vs_oracle or arbitrary_edge or structure_only or masked_edges or neighbors_bit_exact or segment_softmax or fused_gru or shared_plan" \
   > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck.log | tail -3
$T 600 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 120 \
   python -m pytest tests/test_gpu_ba.py -q -m gpu -p no:cacheprovider -k "(ba_vs_oracle and 4-24) or (arbitrary_edge and 0)" \
   > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | tail -3
# fused update operator (tcgen05 / TMA / 2-CTA clusters with DSMEM) and the edge-sharded BA
$T 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 120 \
   python -m pytest tests/test_gpu_gru_mma.py tests/test_gpu_ba.py -q -m gpu -p no:cacheprovider \
   -k "(matches_cublas_path and (3-5 or 4-24)) or deterministic or state or (simulated_ranks and 5-7)" \
   > gpurun_out/sanitize_memcheck_gru.log 2>&1
echo "memcheck(gru) exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_gru.log | tail -3
$T 600 compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 120 \
   python -m pytest tests/test_gpu_gru_mma.py -q -m gpu -p no:cacheprovider -k "matches_cublas_path and 3-5" \
   > gpurun_out/sanitize_racecheck_gru.log 2>&1
echo "racecheck(gru) exit $?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck_gru.log | tail -3
