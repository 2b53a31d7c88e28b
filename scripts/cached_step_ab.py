"""KV-cached rollout engine: tcgen05 GEMMs + separate norms (sea_temporal_small_m(0)) against the small-M fused-norm GEMM
path (1, default): time per model step and parity between the two and against the prefix-recompute plan.
    python scripts/cached_step_ab.py > gpurun_out/cached_step_ab.md"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200._lib import lib  # noqa: E402
from sea_b200.rollout import rollout  # noqa: E402
from sea_b200.temporal import TemporalModel  # noqa: E402

dev = torch.device("cuda", 0)
R = 100


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


print("| config | B | path | ms per 100-step rollout | us per model step | traj-steps/s | rel vs tcgen05 path | rel vs prefix plan |")
print("|---|---:|---|---:|---:|---:|---:|---:|")
for name, E, ln in (("cylinder_flow", 1024, "adaln"), ("multiphase_flow", 2048, "ln")):
    torch.manual_seed(42)
    m = TemporalModel(1, E, 8, 2024, 8, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln).to(dev).eval()
    for B in (32, 8):
        g = torch.Generator(device=dev).manual_seed(1234)
        x0 = torch.randn(B, 1, 2, E, device=dev, generator=g)
        ib = torch.rand(B, 1, 1, device=dev, generator=g).expand(B, R, 1).contiguous()
        prefix = rollout(m, x0, ib, R).clone()
        outs = {}
        for mode in (0, 1):
            lib.sea_temporal_small_m(mode)
            m.engine().__dict__.pop("_cached_plans", None)       # the graphs bake the kernel choice in
            outs[mode] = rollout(m, x0, ib, R, cached=True).clone()
            ms = timed(lambda: rollout(m, x0, ib, R, cached=True, _view_ok=True))
            rel = lambda a, b: ((a - b).norm() / b.norm()).item()
            print(f"| {name} | {B} | {'small-M fused' if mode else 'tcgen05 + norms'} | {ms:.2f} | {ms * 10:.1f} | "
                  f"{B * R / ms * 1e3:.0f} | {rel(outs[mode], outs[0]):.1e} | {rel(outs[mode], prefix):.1e} |", flush=True)
        lib.sea_temporal_small_m(1)
    del m
    torch.cuda.empty_cache()
