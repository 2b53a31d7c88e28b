#!/bin/bash
# round 2, call m: profile evidence of the final kernels: launch list of one rollout, GEMM DRAM traffic over a whole rollout,
# ncu --set full of the GEMM kernel (one mid-size prefix) and of the wide attention backward
mkdir -p gpurun_out
python scripts/rollout_one.py > gpurun_out/r2m_rollout_one.log 2>&1 || exit 1
tail -2 gpurun_out/r2m_rollout_one.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2m_launches_rollout.csv python scripts/rollout_one.py > gpurun_out/r2m_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_bf16 --csv --log-file gpurun_out/r2m_gemm_dram.csv python scripts/rollout_one.py > gpurun_out/r2m_ncu2.log 2>&1
echo "dram pass rc=$?"
python scripts/attn_one.py 4 2024 128 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc2 -c 2 --launch-skip 2 -o gpurun_out/r2m_attn_bwd_wide python scripts/attn_one.py 4 2024 128 > gpurun_out/r2m_ncu3.log 2>&1
echo "attn full rc=$?"
python scripts/profile_forward.py --B 32 --T 100 --reps 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 13 --launch-skip 26 -o gpurun_out/r2m_gemm_t100 python scripts/profile_forward.py --B 32 --T 100 --reps 3 > gpurun_out/r2m_ncu4.log 2>&1
echo "gemm full rc=$?"
ls -la gpurun_out/r2m_*
