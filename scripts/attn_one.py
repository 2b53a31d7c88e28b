"""One fused causal attention forward + backward at max_len (for ncu captures).
    python scripts/attn_one.py [B] [T] [hd]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2024
hd = int(sys.argv[3]) if len(sys.argv) > 3 else 128
nh = 8
dev = torch.device("cuda", 0)
torch.manual_seed(0)
qkv = torch.randn(B * T, 3 * nh * hd, device=dev).bfloat16()
q, k, v = qkv[:, : nh * hd], qkv[:, nh * hd: 2 * nh * hd], qkv[:, 2 * nh * hd:]
for _ in range(2):
    o, lse = ops.attention_fwd(q, k, v, nh, B=B, want_lse=True)
    do = torch.randn_like(o)
    ops.attention_bwd(q, k, v, o, do, lse, nh, B=B)
torch.cuda.synchronize()
print("ok")
