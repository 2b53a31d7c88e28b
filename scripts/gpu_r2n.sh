#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for what in codec attn patch; do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_small.py $what > gpurun_out/r2n_${tool}_${what}.log 2>&1
    echo "$tool $what rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok$' gpurun_out/r2n_${tool}_${what}.log | tr '\n' ' ')"
  done
done
