#!/bin/bash
# round 2, final multi-GPU verification: the data-parallel parity tests (need >= 2 GPUs) and the bench at N = number of GPUs
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -q -m gpu 2>&1 | tail -4 > gpurun_out/r2x_tests_${N}gpu.txt
cat gpurun_out/r2x_tests_${N}gpu.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2x_bench_${N}gpu.json 2> gpurun_out/r2x_bench_${N}gpu.err
echo "bench rc=$?"
grep '^{' gpurun_out/r2x_bench_${N}gpu.json | tail -1 | cut -c1-300
