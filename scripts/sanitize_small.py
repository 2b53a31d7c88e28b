"""Small invocations of the kernels written / reworked in round 2 (ragged / odd shapes, finite-output checks).  Written as a
compute-sanitizer target (`compute-sanitizer --tool memcheck python scripts/sanitize_small.py`); compute-sanitizer is
closed on this GPU pool, so it only runs plainly here: `python scripts/sanitize_small.py [codec|attn|patch]`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200 import ops  # noqa: E402
from sea_b200.patchify import DataPartitioner3D  # noqa: E402
from sea_b200.pipeline import ResidentPipeline  # noqa: E402
from sea_b200.spatial import SpatialModel  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("codec", "all"):
    for C_, D, Hs in ((64, 16, 480), (37, 16, 480), (64, 32, 624)):
        m = SpatialModel([[0, 1], [2]], C_, Hs, 2, D, 8, 2024, 0, 0.0, False, precision="bf16").to(dev).eval()
        x = torch.randn(5, 64, 3, C_, device=dev)
        with torch.no_grad():
            z = m.encode(x)
            y = m.decode(z)
        torch.cuda.synchronize()
        assert torch.isfinite(z).all() and torch.isfinite(y).all()
    print("codec ok")
if what in ("attn", "all"):
    nh = 2
    for hd in (64, 128, 256):
        B, T = 2, 300
        qkv = torch.randn(B * T, 3 * nh * hd, device=dev).bfloat16()
        q, k, v = qkv[:, : nh * hd], qkv[:, nh * hd: 2 * nh * hd], qkv[:, 2 * nh * hd:]
        o, lse = ops.attention_fwd(q, k, v, nh, B=B, want_lse=True)
        do = torch.randn_like(o)
        dq, dk, dv = ops.attention_bwd(q, k, v, o, do, lse, nh, B=B)
        torch.cuda.synchronize()
        assert torch.isfinite(dq.float()).all() and torch.isfinite(dk.float()).all() and torch.isfinite(dv.float()).all()
    print("attention ok")
if what in ("patch", "all"):
    N = 3000
    x, y, z = torch.rand(N), torch.rand(N), torch.rand(N)
    part = DataPartitioner3D(x, y, z, [torch.randn(3, N) for _ in range(2)], m=4, n=5, k=3, device=dev)
    padded, _ = part.create_partitions()
    part.inverse_partition(padded)
    sp = SpatialModel([[0, 1], [2]], 64, 48, 1, 16, 8, 64, 0).to(dev)
    pipe = ResidentPipeline(None, sp, torch.rand(2000), torch.rand(2000), [[0, 1], [2]], feature_range=(-1, 1), device=dev)
    f = torch.randn(3, 2000, 3, device=dev)
    pipe.fit_scalers(f)
    pipe.unpatchify(pipe.patchify(f))
    torch.cuda.synchronize()
    print("patchify ok")
