"""Where does a backward tile's time go?  hd in {64,128}, T = 2024: normal / TMEM-only / barriers-only compute warps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200 import ops  # noqa: E402
from sea_b200._lib import lib  # noqa: E402

dev = torch.device("cuda", 0)
nh = 8
for hd in (128, 64, 256):
    B, T = 4, 2024
    qkv = torch.randn(B * T, 3 * nh * hd, device=dev).bfloat16()
    q, k, v = qkv[:, : nh * hd], qkv[:, nh * hd: 2 * nh * hd], qkv[:, 2 * nh * hd:]
    o, lse = ops.attention_fwd(q, k, v, nh, B=B, want_lse=True)
    do = torch.randn_like(o)
    for mode in (0, 2, 1):
        lib.sea_attention_bwd_probe(mode)
        for _ in range(3):
            ops.attention_bwd(q, k, v, o, do, lse, nh, B=B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.attention_bwd(q, k, v, o, do, lse, nh, B=B)
        e1.record()
        torch.cuda.synchronize()
        print(f"hd {hd} probe {mode}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
    lib.sea_attention_bwd_probe(0)
