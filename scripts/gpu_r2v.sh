#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -5 > gpurun_out/r2v_tests_attn.txt
cat gpurun_out/r2v_tests_attn.txt
timeout 1200 python -m pytest tests/test_temporal_gpu.py tests/test_backward_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2v_tests_model.txt
cat gpurun_out/r2v_tests_model.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
tail -c 600 gpurun_out/r2v_bench.json
