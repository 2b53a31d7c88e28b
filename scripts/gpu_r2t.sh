#!/bin/bash
# epilogue latency work: GEMM parity, then the headline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_temporal_gpu.py tests/test_backward_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2t_tests.txt
cat gpurun_out/r2t_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
tail -c 3000 gpurun_out/r2t_bench.json
