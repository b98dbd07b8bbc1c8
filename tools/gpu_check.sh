#!/bin/bash
# Runs on the B200 box under gpurun: GPU parity tests (risky fast-path tests in their own process so a
# hang cannot hide the other results), smoke, a short bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
T="timeout -k 10"
$T 600 python -m pytest tests/test_gpu_lie.py tests/test_gpu_ba.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_lie_ba.log 2>&1
echo "lie_ba exit $?" >> gpurun_out/summary.txt
$T 600 python -m pytest tests/test_gpu_corr.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_corr.log 2>&1
echo "corr exit $?" >> gpurun_out/summary.txt
$T 600 python -m pytest tests/test_parity_vs_reference_ext.py -q -m gpu -s -p no:cacheprovider > gpurun_out/pytest_refext.log 2>&1
echo "refext exit $?" >> gpurun_out/summary.txt
$T 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/summary.txt
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
$T 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/pytest_lie_ba.log gpurun_out/pytest_corr.log gpurun_out/pytest_refext.log gpurun_out/pytest_engine.log gpurun_out/smoke.log
tail -c 3000 gpurun_out/bench.log
