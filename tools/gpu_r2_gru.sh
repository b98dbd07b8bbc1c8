#!/bin/bash
# GPU check of the rewritten fused update operator: GRU unit tests first (short timeout: a protocol bug shows up as a hang)
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_gru_mma.py -m gpu -q -x 2>&1 | tail -40 | tee gpurun_out/r2_gru_tests.txt
if grep -q "passed" gpurun_out/r2_gru_tests.txt && ! grep -q "failed" gpurun_out/r2_gru_tests.txt; then
  timeout -k 5 600 python -m pytest tests/test_gpu_reference_callers.py tests/test_gpu_engine.py tests/test_gpu_loops.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r2_gru_tests2.txt
  timeout -k 5 300 python bench.py --steps 50 --warmup 5 2>&1 | tail -5 | tee gpurun_out/r2_bench_gru.txt
fi
