#!/bin/bash
# round 2, call b: full GPU test suite + bench after the fp32 backward and the DP bucket work
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/r2b_tests.log 2>&1
tail -5 gpurun_out/r2b_tests.log
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 600 gpurun_out/r2b_bench.json
