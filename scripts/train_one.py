"""Three fwd+bwd+AdamW steps (the last one is steady state: optimizer state exists) of one config (for ncu launch lists of the train step).
    python scripts/train_one.py [cylinder_flow|multiphase_flow] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from sea_b200.optim import AdamW  # noqa: E402
from sea_b200.temporal import TemporalModel  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "multiphase_flow"
E, ln, T, b = (1024, "adaln", 399, 2) if cfg == "cylinder_flow" else (2048, "ln", 199, 4)
if len(sys.argv) > 2:
    b = int(sys.argv[2])
dev = torch.device("cuda", 0)
torch.manual_seed(42)
m = TemporalModel(1, E, 8, 2024, 8, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln).to(dev).train()
opt = AdamW(m.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, engine=m.engine())
x = torch.randn(b, T, 2, E, device=dev)
ib = torch.rand(b, 1, 1, device=dev).expand(b, T, 1).contiguous()
tgt = torch.randn(b, T, 2, E, device=dev)
for _ in range(3):
    opt.zero_grad(set_to_none=True)
    F.mse_loss(m(x, ib), tgt).backward()
    opt.step()
torch.cuda.synchronize()
print("ok")
