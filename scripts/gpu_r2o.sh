#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/rollout_splits.py > gpurun_out/r2o_rollout_splits.md 2> gpurun_out/r2o.err
cat gpurun_out/r2o_rollout_splits.md; tail -5 gpurun_out/r2o.err
