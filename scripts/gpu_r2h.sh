#!/bin/bash
# round 2, call h: two-stream forward schedule: parity + A/B on the headline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_temporal_gpu.py tests/test_dropin_gpu.py tests/test_pipeline_gpu.py -q -x > gpurun_out/r2h_tests.log 2>&1
tail -5 gpurun_out/r2h_tests.log
SEA_TWO_STREAMS=0 python bench.py --no-aux --no-cpu-baseline > gpurun_out/r2h_bench_one_stream.json 2> gpurun_out/r2h_bench.err
python bench.py --no-aux --no-cpu-baseline > gpurun_out/r2h_bench_two_streams.json 2>> gpurun_out/r2h_bench.err
python - <<'PY'
import json
for n in ("one_stream", "two_streams"):
    d = json.load(open(f"gpurun_out/r2h_bench_{n}.json"))
    print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("cached_rollout", {}).get("us_per_model_step"))
PY
