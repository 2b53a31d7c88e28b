#!/bin/bash
# round 2, call e: tensor-core codec v2 + 3-D partitioner tests
mkdir -p gpurun_out
python -m pytest tests/test_spatial_gpu.py tests/test_patchify_gpu.py tests/test_dropin_gpu.py tests/test_pipeline_gpu.py -q -k "spatial or patchify or pipeline or fields or scaled" > gpurun_out/r2e_tests.log 2>&1
tail -5 gpurun_out/r2e_tests.log
timeout 600 python scripts/sweep.py spatial_tc > gpurun_out/r2e_codec_sweep.md 2>&1
cat gpurun_out/r2e_codec_sweep.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spatial_ -c 2 --launch-skip 4 -o gpurun_out/r2e_codec python scripts/codec_one.py > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log
