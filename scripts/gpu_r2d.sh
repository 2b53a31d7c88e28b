#!/bin/bash
# round 2, call d: tensor-core codec: parity tests, sweep, one ncu --set full capture
mkdir -p gpurun_out
python -m pytest tests/test_spatial_gpu.py tests/test_dropin_gpu.py tests/test_pipeline_gpu.py tests/test_optim_gpu.py -q -k "spatial or pipeline or twin or arbitrary or tensor_core or scaled or fields" > gpurun_out/r2d_tests.log 2>&1
tail -30 gpurun_out/r2d_tests.log
timeout 600 python scripts/sweep.py spatial_tc > gpurun_out/r2d_codec_sweep.md 2>&1
cat gpurun_out/r2d_codec_sweep.md
cat > /tmp/codec_one.py <<'PY'
import torch, sys
sys.path.insert(0, ".")
from sea_b200.spatial import SpatialModel
dev = torch.device("cuda")
torch.manual_seed(42)
m = SpatialModel([[0, 1], [2]], 64, 480, 12, 16, 8, 2024, 0, 0.0, False, precision="bf16").to(dev).eval()
x = torch.randn(4000, 64, 3, 64, device=dev)
with torch.no_grad():
    for _ in range(3):
        z = m.encode(x); y = m.decode(z)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spatial_ -c 2 --launch-skip 4 -o gpurun_out/r2d_codec python /tmp/codec_one.py > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
