#!/bin/bash
# ncu --set full captures (B200_PROFILING.md recipe) of our kernels inside the bench step.
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 ncu --set full --clock-control none --import-source on -k regex:corr_fast_kernel -s 4 -c 1 -f -o gpurun_out/corr_fast \
    python bench.py --steps 3 --warmup 3 --profile > gpurun_out/ncu_corr_run.log 2>&1
echo "ncu corr exit $?"
$T 900 ncu --set full --clock-control none --import-source on -k regex:"ba_accumulate|plan_small|segment_softmax|pyramid_pack|transform_kernel|gmap_pack" -s 24 -c 11 -f -o gpurun_out/ba_kernels \
    python bench.py --steps 3 --warmup 3 --profile > gpurun_out/ncu_ba_run.log 2>&1
echo "ncu others exit $?"
