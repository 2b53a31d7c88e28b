#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/attn_bwd_probe.py > gpurun_out/r2g_probe.txt 2>&1
cat gpurun_out/r2g_probe.txt
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_temporal_gpu.py -q -x -k "attention or golden" > gpurun_out/r2g_tests.log 2>&1
tail -3 gpurun_out/r2g_tests.log
timeout 300 python scripts/attn_bench.py > gpurun_out/r2g_attn_bench.md 2>&1
cat gpurun_out/r2g_attn_bench.md
