#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"ba_accumulate" -s 3 -c 3 -f -o gpurun_out/ba_kernels \
    python bench.py --steps 3 --warmup 3 --profile > gpurun_out/ncu_ba_run.log 2>&1
echo "ncu ba exit $?"
