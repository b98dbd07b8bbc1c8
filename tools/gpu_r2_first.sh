#!/bin/bash
# round 2, first GPU call: UMMA shape/rate probe + the full GPU suite (incl. the reference callers on the shims)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout -k 5 120 tools/bin/umma_probe > gpurun_out/r2_umma_probe.txt 2>&1; echo "probe exit $?" | tee -a gpurun_out/r2_umma_probe.txt
tail -32 gpurun_out/r2_umma_probe.txt
timeout -k 10 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_reference_callers.py 2>&1 | tail -15 | tee gpurun_out/r2_pytest_gpu.txt
timeout -k 10 900 python -m pytest tests/test_gpu_reference_callers.py -m gpu -q 2>&1 | tail -60 | tee gpurun_out/r2_pytest_refcallers.txt
