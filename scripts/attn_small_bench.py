"""Short-sequence attention: attention_small.cu vs the tcgen05 kernel, 40 launches per CUDA graph (PDL chain), us per launch.
    python scripts/attn_small_bench.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200 import ops  # noqa: E402
from sea_b200._lib import lib  # noqa: E402

dev = torch.device("cuda")
B, nh = 32, 8
print("| hd | T | short-sequence kernel us | tcgen05 kernel us |\n|---:|---:|---:|---:|")
for hd in (128, 64):
    for T in (4, 16, 32, 50, 64, 100, 128):
        qkv = torch.randn(B * T, 3 * nh * hd, device=dev).bfloat16()
        q, k, v = qkv[:, :nh * hd], qkv[:, nh * hd:2 * nh * hd], qkv[:, 2 * nh * hd:]
        res = []
        for small in (1, 0):
            lib.sea_attention_small(small)
            ops.attention_fwd(q, k, v, nh, B=B)
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side, capture_error_mode="relaxed"):
                for _ in range(40):
                    ops.attention_fwd(q, k, v, nh, B=B)
            torch.cuda.current_stream().wait_stream(side)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 400 * 1e3)
        lib.sea_attention_small(1)
        print(f"| {hd} | {T} | {res[0]:.2f} | {res[1]:.2f} |")
