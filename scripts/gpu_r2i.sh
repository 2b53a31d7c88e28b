#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_modes_gpu.py -q > gpurun_out/r2i_tests.log 2>&1
grep "\[modes\]\|passed\|failed\|FAILED\|Error" gpurun_out/r2i_tests.log | head -60
