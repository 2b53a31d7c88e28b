"""BASELINE configs[4]: scaling sweep of the temporal stack over token count T and field-stream count V
(E fixed per config), forward and forward+backward, plus the ViT-mesh codec over snapshot batch and cells
per patch.  Prints markdown tables (committed under profiles/).

    python scripts/sweep.py [temporal|spatial|all]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import fwd_flops  # noqa: E402
from sea_b200.temporal import TemporalModel  # noqa: E402

dev = torch.device("cuda")


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def temporal():
    print("| config | V | T | B | M=B*T | fwd ms | fwd TFLOP/s | fwd+bwd ms | fwd+bwd TFLOP/s |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for cfg, E, ln in (("cylinder_flow", 1024, "adaln"), ("multiphase_flow", 2048, "ln")):
        for V in (2, 3, 4):
            torch.manual_seed(42)
            m = TemporalModel(1, E, 8, 2024, 8, 0, V, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln).to(dev)
            for T in (64, 128, 256, 399, 512, 1024, 2024):
                if V > 2 and T not in (128, 399, 2024):
                    continue
                B = max(1, 4096 // T)
                x = torch.randn(B, T, V, E, device=dev)
                ib = torch.rand(B, 1, 1, device=dev).expand(B, T, 1).contiguous()
                tgt = torch.randn_like(x)
                fl = fwd_flops(B, T, E=E, H=8 * E, Dd=E // 2, V=V, adaln=(ln == "adaln"))
                m.eval()
                with torch.no_grad():
                    ms_f = timeit(lambda: m(x, ib))
                m.train()

                def step():
                    for p in m.parameters():
                        p.grad = None
                    torch.nn.functional.mse_loss(m(x, ib), tgt).backward()

                ms_fb = timeit(step, reps=3, warm=2)
                print(f"| {cfg} | {V} | {T} | {B} | {B*T} | {ms_f:.3f} | {fl/ms_f/1e9:.0f} | {ms_fb:.3f} | {3*fl/ms_fb/1e9:.0f} |",
                      flush=True)
                del x, ib, tgt
            del m
            torch.cuda.empty_cache()


def spatial(precision="fp32", cs=(32, 64, 128, 256), batches=(128, 1000, 8000)):
    """HBM roofline of the codec: bytes_spatial(min) = 4*P*F*C in + 4*P*G*D latent out per snapshot."""
    from sea_b200.spatial import SpatialModel
    print(f"\ncodec precision = {precision}")
    print("| config | snapshots | cells/patch C | encode ms | encode snapshots/s | encode GB/s (algorithmic) | encode TFLOP/s | decode ms | decode snapshots/s |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for cfg, D, Hs in (("cylinder_flow", 16, 480), ("multiphase_flow", 32, 624)):
        for C in cs:
            torch.manual_seed(42)
            m = SpatialModel([[0, 1], [2]], C, Hs, 12, D, 8, 2024, 0, 0.0, False, precision=precision).to(dev).eval()
            Es = 2 * D
            enc_flops = 2 * 64 * (3 * C) * Hs + 2 * 64 * Hs * D * 2 + 12 * (24 * 64 * Es * Es + 4 * 64 * 64 * Es)
            enc_bytes = 4 * 64 * 3 * C + 4 * 64 * Es
            for Bs in batches:
                x = torch.randn(Bs, 64, 3, C, device=dev)
                try:
                    with torch.no_grad():
                        z = m.encode(x)
                        ms_e = timeit(lambda: m.encode(x))
                        ms_d = timeit(lambda: m.decode(z))
                except RuntimeError as e:
                    print(f"| {cfg} | {Bs} | {C} | unsupported: {e} | | | |", flush=True)
                    continue
                print(f"| {cfg} | {Bs} | {C} | {ms_e:.3f} | {Bs/ms_e*1e3:.0f} | {Bs*enc_bytes/ms_e/1e6:.1f} | "
                      f"{Bs*enc_flops/ms_e/1e9:.2f} | {ms_d:.3f} | {Bs/ms_d*1e3:.0f} |", flush=True)
                del x, z
            del m


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("temporal", "all"):
        temporal()
    if what in ("spatial", "all"):
        spatial("fp32")
        spatial("bf16")
    if what == "spatial_tc":
        spatial("fp32", cs=(64,), batches=(1000, 8000))
        spatial("bf16", cs=(32, 64, 128, 256), batches=(1000, 8000, 32000))
