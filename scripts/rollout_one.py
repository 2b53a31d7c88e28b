"""Two eager 100-step rollouts of the headline workload (cylinder_flow, 32 trajectories): ncu target for the
per-launch list of ONE rollout (skip the first, count the second).
    python scripts/rollout_one.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200.rollout import rollout  # noqa: E402
from sea_b200.temporal import TemporalModel  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(42)
m = TemporalModel(1, 1024, 8, 2024, 8, 0, 2, 2, 0.1, "sea", "learnable", "mlp", "add", 1, 1, True, "adaln").to(dev).eval()
B, R = 32, 100
g = torch.Generator(device=dev).manual_seed(1234)
x0 = torch.randn(B, 1, 2, 1024, device=dev, generator=g)
ib = torch.rand(B, 1, 1, device=dev, generator=g).expand(B, R, 1).contiguous()
eng = m.engine()
for i in range(2):
    n0 = eng.total_launches
    rollout(m, x0, ib, R, graphs=False)
    torch.cuda.synchronize()
    print("library launches in this rollout:", eng.total_launches - n0)
print("ok")
