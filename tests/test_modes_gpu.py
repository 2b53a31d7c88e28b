"""Non-default block / ib variants of the reference (models/temporal.py:103-120, 197-312) through the module-level CUDA
path (sea_b200.modules): the reference's own TemporalModel is built for every exchange_mode / ib mode, deep-copied, one
copy accelerate()d, and forward, loss and every parameter gradient are compared with the eager fp32 copy on the same GPU."""
import copy

import pytest
import torch
import torch.nn.functional as F

from oracle import ref as oref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not oref.available(), reason="reference not staged (oracle/_ref)")]


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ns():
    return oref.load()


CASES = [
    # exchange, ib_scale, ib_add, after_cross, LN
    ("addition", "mlp", "add", True, "adaln"),
    ("addition", "mlp", "add", True, "ln"),
    ("simple", "mlp", "add", True, "adaln"),
    ("pool", "mlp", "add", True, "adaln"),
    ("pool", "mlp", "add", True, "ln"),
    ("sea", "fourier", "add", True, "adaln"),
    ("sea", "linear", "add", True, "ln"),
    ("sea", "mlp", "none", True, "adaln"),
    ("sea", "mlp", "attention", True, "ln"),
    ("sea", "mlp", "add", False, "adaln"),
    ("sea", "mlp", "concat", False, "ln"),       # internal width E + 64: head dim not a multiple of 32 -> attention stays eager
]


WIDTHS = {"small": (512, 4, 4, 3, 40, 128), "cylinder_flow": (1024, 8, 8, 2, 399, 2024)}    # E, heads, scale_ratio, B, T, max_len


@pytest.mark.parametrize("exchange,ib_scale,ib_add,after,ln,width",
                         [c + ("small",) for c in CASES] + [("addition", "mlp", "add", True, "adaln", "cylinder_flow"),
                                                            ("pool", "mlp", "add", True, "adaln", "cylinder_flow"),
                                                            ("sea", "fourier", "add", True, "adaln", "cylinder_flow")])
def test_non_default_modes_forward_and_gradients(cuda, ns, exchange, ib_scale, ib_add, after, ln, width):
    from sea_b200.temporal import accelerate
    torch.manual_seed(3)
    E, nh, sr, B, T, max_len = WIDTHS[width]
    V = 2
    kw = dict(num_layers=1, embed_dim=E, n_heads=nh, max_len=max_len, scale_ratio=sr, src_len=0, num_variables=V, down_proj=2,
              dropout=0.0, exchange_mode=exchange, pos_encoding_mode="learnable", ib_scale_mode=ib_scale,
              ib_addition_mode=ib_add, ib_mlp_layers=1, ib_num=1, add_info_after_cross=after, LN_type=ln)
    ref = ns.temporal.TemporalModel(**kw).to(cuda).train()
    fast = copy.deepcopy(ref)
    accelerate(fast, precision="bf16")
    assert type(fast) is ns.temporal.TemporalModel
    counts = getattr(fast, "_sea_modules", None)
    assert counts is not None and counts["linear"] > 0 and counts["mlp"] == V, counts
    if ib_add != "concat":
        assert counts["self_attention"] == V
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(B, T, V, E, generator=g).to(cuda)
    ib = torch.rand(B, 1, 1, generator=g).expand(B, T, 1).contiguous().to(cuda)
    tgt = torch.randn(B, T, V, E, generator=g).to(cuda)
    ya, yb = ref(x, ib), fast(x, ib)
    la, lb = F.mse_loss(ya, tgt), F.mse_loss(yb, tgt)
    la.backward()
    lb.backward()
    ms = {}
    if width != "small":      # informational (printed, never asserted): forward + backward on this GPU
        for tag, mdl in (("eager", ref), ("modules", fast)):
            def step():
                mdl.zero_grad(set_to_none=True)
                F.mse_loss(mdl(x, ib), tgt).backward()
            step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms[tag] = e0.elapsed_time(e1) / 3
        la = F.mse_loss(ref(x, ib), tgt)
        lb = F.mse_loss(fast(x, ib), tgt)
        ref.zero_grad(set_to_none=True), fast.zero_grad(set_to_none=True)
        la.backward()
        lb.backward()
    e_fwd = _rel(yb, ya)
    worst, n = 0.0, 0
    for (name, p), (_, q) in zip(ref.named_parameters(), fast.named_parameters()):
        assert (p.grad is None) == (q.grad is None), name
        if p.grad is None or p.grad.norm() < 1e-7 * max(1.0, p.numel() ** 0.5) or name.endswith(".k.bias"):
            continue
        cos = F.cosine_similarity(p.grad.flatten().double(), q.grad.flatten().double(), dim=0).item()
        worst = max(worst, 1.0 - cos)
        n += 1
        assert cos > 0.99 and _rel(q.grad, p.grad) < 0.12, (name, cos, _rel(q.grad, p.grad))
    print(f"\\n[modes] {exchange}/{ib_scale}/{ib_add}/after={after}/{ln}: forward rel {e_fwd:.2e}, loss {la.item():.5f} vs "
          f"{lb.item():.5f}, worst 1-cos over {n} gradients {worst:.1e}; switched {counts}"
          + (f"; fwd+bwd {ms['eager']:.2f} ms eager fp32 vs {ms['modules']:.2f} ms module path" if ms else ""))
    assert e_fwd < 2e-2 and abs(la.item() - lb.item()) / la.item() < 1e-2


def test_default_mode_still_uses_the_fused_executor(cuda, ns):
    from sea_b200.temporal import accelerate
    m = ns.temporal.TemporalModel(1, 256, 2, 64, 4, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, "adaln").to(cuda)
    accelerate(m)
    assert hasattr(m, "_sea_engine") and not hasattr(m, "_sea_modules")


@pytest.mark.parametrize("name", ["cylinder_flow", "multiphase_flow"])
def test_encoder_decoder_training_through_the_module_path(cuda, ns, name):
    """Encoder / decoder TRAINING (train/train_encoder.py:205-214, optional in SURVEY §8a): the reference's own
    SpatialModel, built by ProcessData.initialize_spatial_model, with its nn.Linear / MLP / LayerNorm leaves rebound to
    the tcgen05 GEMM and row-norm kernels (forward and backward); the 64-token attention core (head dim 8 / 16) stays
    the reference's eager code.  Loss, every gradient and ten AdamW iterations against the eager fp32 copy."""
    from sea_b200.modules import accelerate_modules
    cfg = oref.temporal_config(name)
    cfg["device"] = str(cuda)
    cfg["dropout_spatial"] = 0.0
    torch.manual_seed(3)
    proc = ns.data_processors.ProcessData(64, cfg)
    eager = proc.initialize_spatial_model().train()
    fast = copy.deepcopy(eager)
    counts = accelerate_modules(fast)
    assert counts["linear"] > 0 and counts["mlp"] > 0, counts
    n_fields = max(max(g) for g in cfg["field_groups"]) + 1
    g = torch.Generator().manual_seed(11)
    x = torch.randn(24, 64, n_fields, 64, generator=g).to(cuda)
    opts = [torch.optim.AdamW(m.parameters(), lr=1e-4) for m in (eager, fast)]
    curves = ([], [])
    for it in range(10):
        for k, (m, opt) in enumerate(zip((eager, fast), opts)):
            opt.zero_grad()
            loss = F.mse_loss(m(x.clone()), x)
            loss.backward()
            curves[k].append(loss.item())
            if it == 0 and k == 1:
                worst = 0.0
                for (pn, pe), (_, pf) in zip(eager.named_parameters(), fast.named_parameters()):
                    if pe.grad is None or pe.grad.norm() < 1e-10:
                        continue
                    assert pf.grad is not None, pn
                    worst = max(worst, _rel(pf.grad, pe.grad))
                print(f"\n[encoder training, module path] {name}: bound {counts}, loss rel {abs(curves[1][0] - curves[0][0]) / curves[0][0]:.2e}, "
                      f"worst gradient rel {worst:.2e}")
                assert worst < 8e-2
        for opt in opts:
            opt.step()
    for a, b in zip(*curves):
        assert abs(a - b) / a < 2e-2
    # informational: one training iteration, eager fp32 vs module path, same GPU
    times = []
    for m, opt in zip((eager, fast), opts):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            opt.zero_grad()
            F.mse_loss(m(x.clone()), x).backward()
            opt.step()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 10)
    print(f"[encoder training, module path] {name}: {times[0]:.2f} ms eager fp32 -> {times[1]:.2f} ms per iteration (24 snapshots)")
