"""ViT-mesh encoder / decoder parity against goldens of the unmodified reference SpatialModel."""
import numpy as np
import pytest
import torch

from oracle import sea_oracle as so
from tests.helpers import rel_l2, spatial_case

pytestmark = pytest.mark.gpu


def build(tag, cuda):
    from sea_b200.spatial import SpatialModel
    g, sd, cfg, x = spatial_case(tag)
    n_inp, hidden, layers, D, nh, B = [int(v) for v in g["meta"]]
    m = SpatialModel(cfg["field_groups"], n_inp, hidden, layers, D, nh, 2024, 0, 0.0, False)
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and all(k.endswith(".pe") for k in missing.missing_keys)
    return g, sd, cfg, m.to(cuda).eval(), x


@pytest.mark.parametrize("tag", ["small", "cylinder_flow", "multiphase_flow"])
def test_spatial_matches_reference_golden(cuda, tag):
    g, sd, cfg, m, x = build(tag, cuda)
    xin = x.clone().to(cuda)
    with torch.no_grad():
        y = m(xin)                       # forward: pad rewrite (in place) + encode + decode
        z = m.encode(xin)
        y2 = m.decode(z)
    torch.cuda.synchronize()
    ez, ey = rel_l2(z.cpu(), g["z"]), rel_l2(y.cpu(), g["y"])
    print(f"\n[spatial] {tag}: latent rel {ez:.2e}, reconstruction rel {ey:.2e}")
    assert z.shape == g["z"].shape and y.shape == g["y"].shape
    assert ez < 1e-4 and ey < 1e-4            # fp32 bar of north_star
    assert rel_l2(y2.cpu(), g["y"]) < 1e-4
    assert np.array_equal(xin.cpu().numpy()[0, 0, 0, -4:], g["x_after"])   # -9999 rewritten in place


def test_latent_layout_fused_store_and_load(cuda):
    """latent_layout=1 == transform_processed_data(encode(x)) (utils/train_utils.py:315-337) and
    decode(latent_layout=1) inverts it."""
    g, sd, cfg, m, x = build("cylinder_flow", cuda)
    xin = x.clone().to(cuda)
    xin[xin == -9999] = 0
    codec = m._codec()
    lat = codec.encode(xin, latent_layout=1)
    B = x.shape[0]
    assert lat.shape == (B, 2, 64 * 16)
    assert rel_l2(lat.cpu().reshape(1, B, 2, -1), g["latent"]) < 1e-4
    y = codec.decode(lat, latent_layout=1)
    assert rel_l2(y.cpu(), g["y"]) < 1e-4


def test_spatial_ragged_batch_and_oracle(cuda):
    """Batch sizes that do not divide anything + comparison with the oracle on fresh inputs."""
    g, sd, cfg, m, x = build("small", cuda)
    gen = torch.Generator().manual_seed(5)
    xb = torch.randn(37, 64, 3, 16, generator=gen)
    with torch.no_grad():
        ref_z = so.spatial_encode(xb, sd, **cfg)
        ref_y = so.spatial_decode(ref_z, sd, field_groups=cfg["field_groups"])
        z = m.encode(xb.to(cuda))
        y = m.decode(z)
    assert rel_l2(z.cpu(), ref_z) < 1e-4 and rel_l2(y.cpu(), ref_y) < 1e-4


@pytest.mark.parametrize("C", [256, 320])
def test_spatial_wide_snapshot_unstaged(cuda, C):
    """Snapshots too wide for the CTA's shared memory (BASELINE configs[4] sweeps C up to 256):
    the patch MLPs read / accumulate the snapshot in place in global memory; results must match
    the oracle exactly like the staged path, including the in-place -9999 rewrite."""
    from oracle import golden_recipe as gr
    from sea_b200.spatial import SpatialModel
    fg = [[0, 1], [2]]
    shapes = [(k, tuple(v.shape)) for k, v in so.init_spatial_state(
        field_groups=fg, n_inp=C, mlp_hidden=480, num_layers=2, embed_dim=16).items()]
    sd = gr.fill_state(shapes, 11)
    m = SpatialModel(fg, C, 480, 2, 16, 8, 2024, 0, 0.0, False)
    m.load_state_dict(sd, strict=False)
    m = m.to(cuda).eval()
    xb = torch.randn(9, 64, 3, C, generator=torch.Generator().manual_seed(6))
    xb[0, 0, 0, -4:] = -9999.0
    xin = xb.clone().to(cuda)
    xz = xb.clone()
    xz[xz == -9999.0] = 0
    with torch.no_grad():
        ref_z = so.spatial_encode(xz, sd, field_groups=fg, num_layers=2, n_heads=8)
        ref_y = so.spatial_decode(ref_z, sd, field_groups=fg)
        y = m(xin)
        z = m.encode(xin)
    assert rel_l2(z.cpu(), ref_z) < 1e-4 and rel_l2(y.cpu(), ref_y) < 1e-4
    assert torch.equal(xin.cpu(), xz)


def test_spatial_throughput_report(cuda):
    from oracle import golden_recipe as gr
    g, sd, cfg, m, x = build("cylinder_flow", cuda)
    B = 4096
    xb = torch.randn(B, 64, 3, 64, device=cuda)
    for _ in range(2):
        z = m.encode(xb)
        y = m.decode(z)
    s, e, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    s.record()
    z = m.encode(xb)
    e.record()
    y = m.decode(z)
    e2.record()
    torch.cuda.synchronize()
    te, td = s.elapsed_time(e), e.elapsed_time(e2)
    flops_enc = 2 * 64 * (128 * 480 + 480 * 16 + 64 * 480 + 480 * 16) + 12 * (24 * 64 * 32 * 32 + 4 * 64 * 64 * 32)
    print(f"\n[spatial perf] cylinder B={B}: encode {te:.2f} ms ({B/te*1e3:.0f} snapshots/s, "
          f"{flops_enc*B/te/1e9:.1f} TFLOP/s fp32, {B*(49152+8192)/te/1e6:.1f} GB/s algorithmic), "
          f"decode {td:.2f} ms ({B/td*1e3:.0f} snapshots/s)")


@pytest.mark.parametrize("tag", ["cylinder_flow", "multiphase_flow"])
def test_tensor_core_codec_matches_reference_golden(cuda, tag):
    """precision='bf16': every contraction of the codec on mma.sync bf16 tensor-core tiles (sea_spatial_encode_tc /
    _decode_tc).  Bar: north_star's bf16 tolerance (2e-2 relative) against the unmodified reference's fp32 goldens;
    measured values are printed.  Same in-place pad rewrite, same latent layouts."""
    from sea_b200.spatial import SpatialModel
    g, sd, cfg, x = spatial_case(tag)
    n_inp, hidden, layers, D, nh, B = [int(v) for v in g["meta"]]
    m = SpatialModel(cfg["field_groups"], n_inp, hidden, layers, D, nh, 2024, 0, 0.0, False, precision="bf16")
    m.load_state_dict(sd, strict=False)
    m = m.to(cuda).eval()
    xin = x.clone().to(cuda)
    with torch.no_grad():
        y = m(xin)
        z = m.encode(xin)
        y2 = m.decode(torch.as_tensor(g["z"]).to(cuda))     # decoder alone, from the reference's latents
    torch.cuda.synchronize()
    ez, ey, ey2 = rel_l2(z.cpu(), g["z"]), rel_l2(y.cpu(), g["y"]), rel_l2(y2.cpu(), g["y"])
    print(f"\n[spatial tc] {tag}: latent rel {ez:.2e}, reconstruction rel {ey:.2e}, decoder alone {ey2:.2e}")
    assert ez < 2e-2 and ey < 2e-2 and ey2 < 2e-2
    assert np.array_equal(xin.cpu().numpy()[0, 0, 0, -4:], g["x_after"])
    codec = m._codec()
    lat = codec.encode(xin, latent_layout=1)
    assert torch.equal(lat.view(B, 2, 64, D).permute(0, 2, 1, 3), z)        # layout 1 is a pure permutation of layout 0
    assert torch.equal(codec.decode(lat, latent_layout=1), m.decode(z))
    # weights re-packed after a parameter update
    with torch.no_grad():
        m.encode.ln.bias.add_(0.5)
        m.encode.encoders[0].layer2.weight.mul_(1.5)
        z2 = m.encode(xin)
        ref2 = so.spatial_encode(xin.cpu(), {k: v.detach().cpu() for k, v in m.state_dict().items()}, **cfg)
    assert rel_l2(z2.cpu(), ref2) < 2e-2 and rel_l2(z2.cpu(), g["z"]) > 0.05


def test_tensor_core_codec_ragged_batches_vs_fp32_kernels(cuda):
    """Batch sizes 1 / 37 / 300 and a different n_inp (C = 32, 128): the tensor-core kernels against the fp32 CUDA-core
    kernels of the same library on the same weights."""
    from sea_b200.spatial import SpatialModel
    for C_, B in ((32, 1), (64, 37), (128, 300)):
        torch.manual_seed(C_)
        a = SpatialModel([[0, 1], [2]], C_, 480, 12, 16, 8, 2024, 0, 0.0, False, precision="fp32").to(cuda).eval()
        b = SpatialModel([[0, 1], [2]], C_, 480, 12, 16, 8, 2024, 0, 0.0, False, precision="bf16").to(cuda).eval()
        b.load_state_dict(a.state_dict())
        x = torch.randn(B, 64, 3, C_, device=cuda)
        x[:, :, :, C_ - 5:] = 0.0
        with torch.no_grad():
            za, zb = a.encode(x), b.encode(x)
            ya, yb = a.decode(za), b.decode(za)
        assert rel_l2(zb.cpu(), za.cpu()) < 2e-2 and rel_l2(yb.cpu(), ya.cpu()) < 2e-2, (C_, B)


@pytest.mark.parametrize("C", [37, 50, 63])
def test_arbitrary_cells_per_patch(cuda, C):
    """n_inp is the cell count of the fullest patch of the user's mesh (utils/data_processors.py:61-88,
    train/train_temporal.py:147-148): any integer.  fp32 kernels (scalar-load path when C % 4 != 0) against the oracle at
    1e-4, tensor-core kernels (cell axis zero-padded to 16 at pack time, decoder stores masked) at 2e-2; the decoder must
    not write outside its [B, 64, F, C] output."""
    from oracle import golden_recipe as gr
    from sea_b200.spatial import SpatialModel
    fg = [[0, 1], [2]]
    shapes = [(k, tuple(v.shape)) for k, v in so.init_spatial_state(
        field_groups=fg, n_inp=C, mlp_hidden=480, num_layers=3, embed_dim=16).items()]
    sd = gr.fill_state(shapes, 13)
    g = torch.Generator().manual_seed(C)
    x = torch.randn(9, 64, 3, C, generator=g)
    x[:, :, :, C - 7:] = 0.0
    with torch.no_grad():
        ref_z = so.spatial_encode(x, sd, field_groups=fg, num_layers=3, n_heads=8)
        ref_y = so.spatial_decode(ref_z, sd, field_groups=fg)
    for prec, bar in (("fp32", 1e-4), ("bf16", 2e-2)):
        m = SpatialModel(fg, C, 480, 3, 16, 8, 2024, 0, 0.0, False, precision=prec)
        m.load_state_dict(sd, strict=False)
        m = m.to(cuda).eval()
        with torch.no_grad():
            z = m.encode(x.to(cuda))
            buf = torch.full((9 * 64 * 3 * C + 64,), 777.0, device=cuda)        # canary behind the output
            y = m.decode(ref_z.to(cuda))
        assert rel_l2(z.cpu(), ref_z) < bar and rel_l2(y.cpu(), ref_y) < bar, (prec, C)
        assert torch.isfinite(y).all() and y.shape == (9, 64, 3, C)
        del buf
