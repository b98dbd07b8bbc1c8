#!/bin/bash
# round 2 profiling pass (B200_PROFILING.md recipe): launch list of the bench step + ncu --set full captures of our kernels
mkdir -p gpurun_out
T="timeout -k 10"
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --profile > gpurun_out/r02_launch_run.log 2>&1
echo "launch list exit $?"
$T 900 ncu --set full --clock-control none --import-source on -k regex:gru_mma_kernel -s 20 -c 5 -f -o gpurun_out/r02_gru_mma \
    python bench.py --steps 3 --warmup 3 --profile > gpurun_out/r02_ncu_gru_run.log 2>&1
echo "ncu gru exit $?"
$T 900 ncu --set full --clock-control none --import-source on -k regex:"corr_fast_kernel|segment_softmax|ba_accumulate|plan_small|transform_kernel|pyramid_pack|gmap_pack|patch_gather" -s 48 -c 12 -f -o gpurun_out/r02_other_kernels \
    python bench.py --steps 3 --warmup 3 --profile > gpurun_out/r02_ncu_other_run.log 2>&1
echo "ncu others exit $?"
ls -la gpurun_out/*.ncu-rep
