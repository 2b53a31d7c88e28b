"""Micro-batched rollout plans: headline workload (cylinder_flow, 32 trajectories x 100 steps) and the multiphase one with the
trajectories split into 1 / 2 / 4 groups on their own streams; prefix-recompute plan and KV-cached engine.
    python scripts/rollout_splits.py > gpurun_out/rollout_splits.md"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200.rollout import rollout  # noqa: E402
from sea_b200.temporal import TemporalModel  # noqa: E402

dev = torch.device("cuda", 0)
B, R = 32, 100


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("| config | engine | splits | ms per 100-step rollout | traj-steps/s | max rel diff vs splits = 1 |")
print("|---|---|---:|---:|---:|---:|")
for name, E, ln in (("cylinder_flow", 1024, "adaln"), ("multiphase_flow", 2048, "ln")):
    torch.manual_seed(42)
    m = TemporalModel(1, E, 8, 2024, 8, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1234)
    x0 = torch.randn(B, 1, 2, E, device=dev, generator=g)
    ib = torch.rand(B, 1, 1, device=dev, generator=g).expand(B, R, 1).contiguous()
    for cached in (False, True):
        ref = None
        for splits in (1, 2, 4):
            out = rollout(m, x0, ib, R, cached=cached, splits=splits).clone()
            if ref is None:
                ref = out
            rel = ((out - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)).max().item()
            ms = timed(lambda: rollout(m, x0, ib, R, cached=cached, splits=splits, _view_ok=True))
            print(f"| {name} | {'KV-cached' if cached else 'prefix recompute'} | {splits} | {ms:.2f} | {B * R / ms * 1e3:.0f} | {rel:.1e} |",
                  flush=True)
    del m
    torch.cuda.empty_cache()
