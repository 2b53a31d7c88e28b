#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_dropout_gpu.py tests/test_ops_gpu.py -q -x -k "attention or dropout or gradients" > gpurun_out/r2l_tests.log 2>&1
tail -8 gpurun_out/r2l_tests.log
SEA_BWD_WIDE=1 timeout 300 python scripts/attn_bench.py > gpurun_out/r2l_attn_wide.md 2>&1
SEA_BWD_WIDE=0 timeout 300 python scripts/attn_bench.py > gpurun_out/r2l_attn_narrow.md 2>&1
cat gpurun_out/r2l_attn_wide.md; grep "| 2024 " gpurun_out/r2l_attn_narrow.md
