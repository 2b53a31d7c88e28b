#!/bin/bash
# ncu --set full of the final kernels: the 13 GEMM launches of one forward at B=32, T=100, and the short-sequence attention
mkdir -p gpurun_out
python scripts/profile_forward.py --B 32 --T 100 --reps 3 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 13 --launch-skip 26 -o gpurun_out/r2z_gemm_t100 python scripts/profile_forward.py --B 32 --T 100 --reps 3 > gpurun_out/r2z_ncu1.log 2>&1
echo "gemm full rc=$?"
python scripts/profile_forward.py --B 32 --T 16 --reps 3 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_small -c 3 --launch-skip 6 -o gpurun_out/r2z_attn_small_t16 python scripts/profile_forward.py --B 32 --T 16 --reps 3 > gpurun_out/r2z_ncu2.log 2>&1
echo "attn small full rc=$?"
ls -la gpurun_out/r2z_*
