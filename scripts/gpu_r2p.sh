#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_temporal_gpu.py tests/test_pipeline_gpu.py -q -x -k "cached or rollout or step or pipeline or drift or micro" > gpurun_out/r2p_tests.log 2>&1
tail -6 gpurun_out/r2p_tests.log
timeout 600 python scripts/cached_step_ab.py > gpurun_out/r2p_cached_ab.md 2> gpurun_out/r2p.err
cat gpurun_out/r2p_cached_ab.md; tail -5 gpurun_out/r2p.err
