"""Fused causal attention kernels alone: forward / backward TFLOP/s (causal-useful FLOPs: 2*B*nh*T^2*hd forward, 2.5x
backward) for the three head geometries of the two configs at the configs' own T and at max_len.
    python scripts/attn_bench.py > gpurun_out/attn_bench.md"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
if os.environ.get("SEA_BWD_WIDE") is not None:      # A/B of the two backward plans
    from sea_b200._lib import lib
    lib.sea_attention_bwd_wide(int(os.environ["SEA_BWD_WIDE"]))
    print(f"backward plan: {'128-wide, two-half hand-over' if int(os.environ['SEA_BWD_WIDE']) else '64-wide, double-buffered'}")
pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
sustained = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1396.0)))
nh = 8


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"sustained bf16 peak used as denominator: {sustained:.1f} TFLOP/s\n")
print("| head dim | B | T | fwd us | fwd TFLOP/s | frac | bwd us | bwd TFLOP/s | frac |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for hd in (64, 128, 256):
    for B, T in ((32, 199), (16, 399), (4, 2024)):
        g = torch.Generator(device=dev).manual_seed(9)
        qkv = torch.randn(B * T, 3 * nh * hd, device=dev, generator=g).bfloat16()
        q, k, v = qkv[:, : nh * hd], qkv[:, nh * hd: 2 * nh * hd], qkv[:, 2 * nh * hd:]
        o, lse = ops.attention_fwd(q, k, v, nh, B=B, want_lse=True)
        do = torch.randn(B * T, nh * hd, device=dev, generator=g).bfloat16()
        ms_f = timed(lambda: ops.attention_fwd(q, k, v, nh, B=B, want_lse=True), 30)
        ms_b = timed(lambda: ops.attention_bwd(q, k, v, o, do, lse, nh, B=B), 20)
        fl = 2.0 * B * nh * T * T * hd
        tf_f, tf_b = fl / (ms_f * 1e-3) / 1e12, 2.5 * fl / (ms_b * 1e-3) / 1e12
        print(f"| {hd} | {B} | {T} | {ms_f*1e3:.1f} | {tf_f:.0f} | {tf_f/sustained:.2f} | {ms_b*1e3:.1f} | {tf_b:.0f} | "
              f"{tf_b/sustained:.2f} |", flush=True)
