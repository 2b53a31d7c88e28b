#!/bin/bash
# round 2, call k (8 GPUs): the driver's scaling command at N = 8
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2k_bench_8gpu.json 2> gpurun_out/r2k_bench_8gpu.err
tail -c 300 gpurun_out/r2k_bench_8gpu.json; tail -5 gpurun_out/r2k_bench_8gpu.err
