#!/bin/bash
# round 2, call j: full GPU suite + full bench on one GPU
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; tail -2 gpurun_out/r2j_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/r2j_tests.log 2>&1
tail -6 gpurun_out/r2j_tests.log
python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -c 400 gpurun_out/r2j_bench.json; tail -3 gpurun_out/r2j_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2j_bench_reference.json 2>> gpurun_out/r2j_bench.err
cat gpurun_out/r2j_bench_reference.json | cut -c1-300
