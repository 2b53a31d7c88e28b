#!/bin/bash
# tensor-pipe / SFU / issue utilisation of the attention kernels, all three head dims, T = 2024 (one metrics pass each)
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum
for hd in 64 128 256; do
  python scripts/attn_one.py 4 2024 $hd > /dev/null 2>&1 || exit 1
  ncu --metrics $M --clock-control none -k regex:"attn_fwd|attn_bwd_tc" --launch-skip 3 -c 3 --csv --log-file gpurun_out/r2r_attn_hd$hd.csv python scripts/attn_one.py 4 2024 $hd > /dev/null 2>&1
  echo "hd $hd rc=$?"
done
ls -la gpurun_out/r2r_*
