#!/bin/bash
# round 2, call f: pipelined attention backward (parity + speed), codec v3 (3 CTAs / SM)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_dropout_gpu.py tests/test_spatial_gpu.py -q -x -k "attention or spatial or tensor_core or arbitrary or dropout" > gpurun_out/r2f_tests.log 2>&1
tail -15 gpurun_out/r2f_tests.log
timeout 300 python scripts/attn_bench.py > gpurun_out/r2f_attn_bench.md 2>&1
cat gpurun_out/r2f_attn_bench.md
timeout 600 python scripts/sweep.py spatial_tc > gpurun_out/r2f_codec_sweep.md 2>&1
grep "| 8000 \|precision" gpurun_out/r2f_codec_sweep.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -c 2 --launch-skip 2 -o gpurun_out/r2f_attn_bwd python scripts/attn_one.py 4 2024 128 > gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_ncu.log
