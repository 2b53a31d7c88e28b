import sys, torch
sys.path.insert(0, "/root/repo")
from sea_b200 import ops
dev = torch.device("cuda")
for M, H in ((3184, 16384), (796, 16384), (6384, 8192)):
    h = (torch.randn(M, H, device=dev) * 1.5).bfloat16(); dg = torch.randn(M, H, device=dev).bfloat16()
    w = torch.ones(H, device=dev); b = torch.zeros(H, device=dev)
    _, st = ops.ln_gelu_fwd_with_stats(h, w, b)
    for _ in range(3): ops.ln_gelu_bwd(dg, h, st, w, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.ln_gelu_bwd(dg, h, st, w, b)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"ln_gelu_bwd M={M} H={H}: {us:.1f} us  {6.0*M*H/us/1e6:.2f} TB/s")
