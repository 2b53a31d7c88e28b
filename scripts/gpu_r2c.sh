#!/bin/bash
# round 2, call c (2 GPUs): DP parity tests, pipeline tests, DP train bench with the bucketed bf16 exchange
mkdir -p gpurun_out
python -m pytest tests/test_multigpu_gpu.py tests/test_pipeline_gpu.py tests/test_optim_gpu.py -q > gpurun_out/r2c_tests.log 2>&1
tail -15 gpurun_out/r2c_tests.log
for cfg in cylinder_flow multiphase_flow; do
  for b in 2 16; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      scripts/dp_train_bench.py --config $cfg --b $b >> gpurun_out/r2c_dp.jsonl 2>> gpurun_out/r2c_dp.err
  done
done
cat gpurun_out/r2c_dp.jsonl
