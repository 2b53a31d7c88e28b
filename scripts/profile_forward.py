"""Small driver for ncu: N forwards of the cylinder_flow temporal model at a fixed (B, T).
    python scripts/profile_forward.py --B 32 --T 81 --reps 3 [--train] [--config multiphase_flow]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200.temporal import TemporalModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--T", type=int, default=81)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--config", default="cylinder_flow")
ap.add_argument("--train", action="store_true")
ap.add_argument("--varying-ib", action="store_true")
a = ap.parse_args()
E, ln = (1024, "adaln") if a.config == "cylinder_flow" else (2048, "ln")
torch.manual_seed(42)
m = TemporalModel(1, E, 8, 2024, 8, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln).cuda()
x = torch.randn(a.B, a.T, 2, E, device="cuda")
ib = torch.rand(a.B, 1, 1, device="cuda").expand(a.B, a.T, 1).contiguous()
if a.varying_ib:
    ib = torch.rand(a.B, a.T, 1, device="cuda")
if a.train:
    m.train()
    tgt = torch.randn_like(x)
    for _ in range(a.reps):
        m.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(m(x, ib), tgt)
        loss.backward()
else:
    m.eval()
    m.engine().ib_time_invariant = not a.varying_ib
    with torch.no_grad():
        for _ in range(a.reps):
            y = m(x, ib)
torch.cuda.synchronize()
print("ok", a)
