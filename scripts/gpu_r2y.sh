#!/bin/bash
# after the register cap of the GEMM kernel: GEMM / op / model / backward / optimiser parity, smoke, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_temporal_gpu.py tests/test_backward_gpu.py tests/test_optim_gpu.py -q -m gpu 2>&1 | tail -4 > gpurun_out/r2y_tests.txt
cat gpurun_out/r2y_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
