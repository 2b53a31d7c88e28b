"""Host-side cost of one cylinder_flow train step (is the step launch-bound?): wall time of the Python call
without synchronisation against the CUDA-event time, plus a cProfile of the host work."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from sea_b200.optim import AdamW  # noqa: E402
from sea_b200.temporal import TemporalModel  # noqa: E402

dev = torch.device("cuda", 0)
cfg = sys.argv[1] if len(sys.argv) > 1 else "cylinder_flow"
E, ln, T, b = (1024, "adaln", 399, 2) if cfg == "cylinder_flow" else (2048, "ln", 199, 4)
torch.manual_seed(42)
m = TemporalModel(1, E, 8, 2024, 8, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln).to(dev).train()
opt = AdamW(m.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, engine=m.engine())
x = torch.randn(b, T, 2, E, device=dev)
ib = torch.rand(b, 1, 1, device=dev).expand(b, T, 1).contiguous()
tgt = torch.randn(b, T, 2, E, device=dev)
eng = m.engine()


def step():
    opt.zero_grad(set_to_none=True)
    F.mse_loss(m(x, ib), tgt).backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
n = 20
l0 = eng.total_launches
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    step()
e1.record()
t_host = (time.perf_counter() - t0) / n * 1e3
torch.cuda.synchronize()
print(f"{cfg}: host {t_host:.3f} ms/step (no sync), CUDA events {e0.elapsed_time(e1)/n:.3f} ms/step, "
      f"{(eng.total_launches - l0)/n:.0f} library launches/step")
# forward-only and backward-only host cost
out = m(x, ib)
loss = F.mse_loss(out, tgt)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    out = m(x, ib)
t_f = (time.perf_counter() - t0) / n * 1e3
torch.cuda.synchronize()
print(f"forward call (host, autograd on): {t_f:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
