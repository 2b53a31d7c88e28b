#!/bin/bash
# Round-2 GPU pass 1: parity suites (incl. the drop-in tests on the staged reference), small-M GEMM latency trace,
# bench line, then ONE ncu pass (DRAM bytes of the GEMM launches of a whole rollout).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -q -x --ignore=tests/test_dropin_gpu.py > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2a_tests.log
python -m pytest tests/test_dropin_gpu.py -m gpu -q -s > gpurun_out/r2a_dropin.log 2>&1; echo "dropin rc=$?"
grep -E "^\[|passed|failed|FAILED|Error" gpurun_out/r2a_dropin.log | head -60
python scripts/gemm_latency.py > gpurun_out/r2a_gemm_latency.txt 2>&1; echo "latency rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.err
python scripts/rollout_one.py > gpurun_out/r2a_rollout_one.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm_bf16 --csv --log-file gpurun_out/r2a_gemm_dram.csv python scripts/rollout_one.py > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?"
