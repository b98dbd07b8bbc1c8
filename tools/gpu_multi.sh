#!/bin/bash
# 2-GPU run of the bench exactly as the driver launches it
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_2gpu.log 2>&1
echo "2gpu exit $?"; tail -c 1500 gpurun_out/bench_2gpu.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_2gpu_ref.log 2>&1
echo "2gpu ref exit $?"; tail -c 600 gpurun_out/bench_2gpu_ref.log
