"""One encode + decode of 4000 cylinder_flow snapshots on the tensor-core codec (ncu capture target)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200.spatial import SpatialModel  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(42)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
m = SpatialModel([[0, 1], [2]], 64, 480, 12, 16, 8, 2024, 0, 0.0, False, precision=prec).to(dev).eval()
x = torch.randn(4000, 64, 3, 64, device=dev)
with torch.no_grad():
    for _ in range(3):
        z = m.encode(x)
        y = m.decode(z)
torch.cuda.synchronize()
