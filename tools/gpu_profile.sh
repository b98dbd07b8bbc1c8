#!/bin/bash
# ncu evidence (B200_PROFILING.md recipe): (1) launch list of the bench step, (2) --set full capture of the
# dominant kernel.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --profile > gpurun_out/launches_run.log 2>&1
echo "launch list exit $?"
$T 900 ncu --set full --clock-control none --import-source on -k regex:corr_fast_kernel -s 4 -c 3 -f -o gpurun_out/corr_fast \
    python bench.py --steps 3 --warmup 3 --profile > gpurun_out/ncu_corr_run.log 2>&1
echo "ncu corr exit $?"
$T 900 ncu --set full --clock-control none --import-source on -k regex:"ba_accumulate|ba_solve|plan_small" -s 6 -c 6 -f -o gpurun_out/ba_kernels \
    python bench.py --steps 3 --warmup 3 --profile > gpurun_out/ncu_ba_run.log 2>&1
echo "ncu ba exit $?"
ls -la gpurun_out | head -30
