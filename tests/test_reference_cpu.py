"""CPU checks around the staged reference (oracle/_ref): it is unmodified, it imports, the oracle restatement agrees
with the LIVE reference classes on fresh seeded inputs (beyond the committed goldens), and the install() hook wraps /
unwraps the reference factories without touching CPU models."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import ref as oref
from oracle import sea_oracle as so

pytestmark = pytest.mark.skipif(not oref.available(), reason="reference not staged (run __graft_entry__.build())")


@pytest.fixture(scope="module")
def ns():
    return oref.load()


def test_staged_reference_is_unmodified():
    if oref.root() != oref.STAGED:
        pytest.skip("running against /root/reference directly")
    assert oref.verify()
    if os.path.isdir(oref.SOURCE):      # build container: byte-identical to the source tree
        import filecmp
        for pkg in oref.PACKAGES:
            cmp = filecmp.dircmp(os.path.join(oref.SOURCE, pkg), os.path.join(oref.STAGED, pkg), ignore=["__pycache__"])
            assert not cmp.diff_files and not cmp.left_only and not cmp.right_only, (pkg, cmp.diff_files)


@pytest.mark.parametrize("ln", ["adaln", "ln"])
def test_oracle_matches_live_reference(ns, ln):
    """forward, loss and every gradient of the oracle restatement vs the reference's own nn.Module on fresh inputs."""
    torch.manual_seed(5)
    E, nh, V, B, T = 64, 2, 3, 2, 11
    m = ns.temporal.TemporalModel(2, E, nh, 32, 2, 0, V, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln)
    with torch.no_grad():   # non-trivial biases / norm parameters
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    x, ib, tgt = torch.randn(B, T, V, E), torch.rand(B, T, 1), torch.randn(B, T, V, E)
    loss_ref = F.mse_loss(m(x, ib), tgt)
    loss_ref.backward()
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items() if v.is_floating_point()}
    y = so.temporal_forward(x, ib, sd, num_layers=2, n_heads=nh, ln_type=ln)
    loss = F.mse_loss(y, tgt)
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-6 * abs(loss_ref.item()) + 1e-9
    n = 0
    for name, p in m.named_parameters():
        if p.grad is None:
            assert sd[name].grad is None or sd[name].grad.abs().max() == 0, name
            continue
        err = (sd[name].grad - p.grad).norm() / p.grad.norm().clamp_min(1e-12)
        # the TIPI input layer sits under a LayerNorm over scale_ratio*ib_num = 2 values: pure cancellation in fp32
        assert err < (2e-3 if ".ib.layers.0." in name else 2e-5), (name, err.item())
        n += 1
    assert n > 100


def test_install_hook_wraps_and_restores(ns):
    import sea_b200
    tt, dp = ns.train_temporal, ns.data_processors
    ref_get_model, ref_init = tt.get_model, dp.ProcessData.initialize_spatial_model
    cfg = oref.temporal_config("cylinder_flow")
    cfg.update(embed_dim=64, n_heads=2, block_size=16, scale_ratio=2, device="cpu")
    sea_b200.install()
    try:
        assert tt.get_model is not ref_get_model and tt.get_model.__wrapped__ is ref_get_model
        assert dp.ProcessData.initialize_spatial_model.__wrapped__ is ref_init
        sea_b200.install(precision="fp32")            # idempotent
        assert tt.get_model.__wrapped__ is ref_get_model
        model, loss_fn, opt = tt.get_model(cfg, torch.device("cpu"))
        # a CPU model is handed back untouched: no engine, the reference's own forward
        assert type(model) is ns.temporal.TemporalModel and not hasattr(model, "_sea_engine")
        assert isinstance(opt, torch.optim.AdamW) and isinstance(loss_fn, torch.nn.MSELoss)
        y = model(torch.randn(1, 4, 2, 64), torch.rand(1, 4, 1))
        assert y.shape == (1, 4, 2, 64)
    finally:
        sea_b200.uninstall()
    assert tt.get_model is ref_get_model and dp.ProcessData.initialize_spatial_model is ref_init


def test_accelerate_on_reference_instance_builds_descriptor_and_refuses_cpu(ns):
    from sea_b200.temporal import accelerate
    m = ns.temporal.TemporalModel(1, 128, 2, 64, 2, 0, 2, 2, 0.1, "sea", "learnable", "mlp", "add", 1, 1, True, "adaln")
    accelerate(m)
    eng = m._sea_engine
    h = eng._hyper()
    assert (h["E"], h["nh"], h["H"], h["Dd"], h["V"], h["kind"]) == (128, 2, 256, 64, 2, "adaln")
    assert eng.dropout == pytest.approx(0.1)
    assert len(eng._live_params()) == sum(1 for _ in m.parameters()) - 34      # 34 dead tensors
    with pytest.raises(RuntimeError, match="no CPU path"):
        with torch.no_grad():
            m(torch.zeros(1, 4, 2, 128), torch.zeros(1, 4, 1))
    # non-default exchange modes go to the module-level path (sea_b200.modules): bf16 only, and CUDA only
    for other in ("pool", "addition"):
        mb = ns.temporal.TemporalModel(1, 128, 2, 64, 2, 0, 2, 2, 0.0, other, "learnable", "mlp", "add", 1, 1, True, "ln")
        with pytest.raises(NotImplementedError):
            accelerate(mb, precision="fp32")
        with pytest.raises(RuntimeError, match="no CPU path"):
            accelerate(mb)
