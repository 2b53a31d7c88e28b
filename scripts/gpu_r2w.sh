#!/bin/bash
# round 2, final verification on one GPU: whole GPU suite, smoke, both bench arms, launch list + GEMM DRAM pass of the final kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r2w_tests.txt
cat gpurun_out/r2w_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2w_smoke.txt
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err; echo "reference arm rc=$?"
python scripts/rollout_one.py > gpurun_out/r2w_rollout_one.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2w_launches_rollout.csv python scripts/rollout_one.py > gpurun_out/r2w_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 --csv --log-file gpurun_out/r2w_gemm_dram.csv python scripts/rollout_one.py > gpurun_out/r2w_ncu2.log 2>&1
echo "dram pass rc=$?"
tail -c 400 gpurun_out/r2w_bench.json; tail -c 300 gpurun_out/r2w_bench_reference.json
