#!/bin/bash
# launch lists of one train step at the reference batch sizes (both configs)
mkdir -p gpurun_out
for cfg in cylinder_flow multiphase_flow; do
  python scripts/train_one.py $cfg > /dev/null 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2s_train_$cfg.csv python scripts/train_one.py $cfg > /dev/null 2>&1
  echo "$cfg rc=$?"
done
