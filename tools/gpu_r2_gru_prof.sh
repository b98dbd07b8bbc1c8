#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gru_timing.py 2>&1 | tee gpurun_out/r2_gru_timeline1.txt
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gru_mma_kernel -s 6 -c 6 -f -o gpurun_out/r2_gru_mma \
    python tools/gru_timing.py > gpurun_out/r2_ncu_gru.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
