"""The drop-in, proven on the reference's OWN classes and entry points (north_star: module classes,
train/train_temporal.py and main.py stay unchanged).

The unmodified reference travels to the GPU box as ``oracle/_ref`` (staged by ``__graft_entry__.build()``,
sha256-verified here).  Every test builds the reference's objects through the reference's own code
(``get_model`` train/train_temporal.py:190-223 with the config dicts of configs/*.py), deep-copies them, lets
one copy run the reference's eager fp32 PyTorch path ON THE SAME GPU and switches the other to the CUDA path
with the hook of INTEGRATION.md (``sea_b200.install()`` / ``accelerate()``), then compares: forward, 10-step
rollout, loss, every ``.grad``, 100 iterations of the reference's loop body with ``torch.optim.AdamW`` at the
config's learning rate, and the reference's rollout loop (``autoregressive_validation``
utils/train_utils.py:154-185) called verbatim.  Shapes are the configs' own: cylinder_flow [2,399,2,1024],
multiphase_flow [4,199,2,2048].
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader

from oracle import ref as oref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not oref.available(), reason="reference not staged (oracle/_ref)")]

CASES = {"cylinder_flow": dict(B=2, T=399), "multiphase_flow": dict(B=4, T=199)}


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ns():
    if oref.root() == oref.STAGED:
        assert oref.verify(), "oracle/_ref differs from the staged manifest (the reference must be unmodified)"
    return oref.load()


def _cfg(name, device, dropout=0.0):
    cfg = oref.temporal_config(name)
    cfg["device"] = str(device)
    cfg["dropout"] = dropout
    return cfg


def _pair(ns, name, device, precision, dropout=0.0, seed=42):
    """(eager reference model, accelerated deep copy, their loss fns and optimizers) built by the reference's
    own get_model — the accelerated one through the installed hook."""
    import sea_b200
    cfg = _cfg(name, device, dropout)
    torch.manual_seed(seed)
    assert not hasattr(ns.train_temporal.get_model, "__wrapped__")       # the reference's own function
    ref_model, ref_loss, ref_opt = ns.train_temporal.get_model(cfg, device)
    sea_b200.install(precision=precision, spatial=False)
    try:
        torch.manual_seed(seed)
        fast_model, fast_loss, fast_opt = ns.train_temporal.get_model(cfg, device)
    finally:
        sea_b200.uninstall()
    assert type(fast_model) is ns.temporal.TemporalModel and hasattr(fast_model, "_sea_engine")
    assert isinstance(fast_opt, torch.optim.AdamW) and type(fast_opt) is type(ref_opt)
    fast_model.load_state_dict(ref_model.state_dict())      # same init (both were seeded, make it explicit)
    return cfg, (ref_model, ref_loss, ref_opt), (fast_model, fast_loss, fast_opt)


def _batch(name, device, seed=7, T=None):
    c = CASES[name]
    E = 1024 if name == "cylinder_flow" else 2048
    B, T = c["B"], (T or c["T"])
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(B, T, 2, E, generator=g)
    ib = torch.rand(B, 1, 1, generator=g).expand(B, T, 1).contiguous()
    tgt = torch.roll(x, -1, dims=1) + 0.1 * torch.randn(B, T, 2, E, generator=g)
    return x.to(device), ib.to(device), tgt.to(device)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("precision,bar", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_accelerated_reference_forward_and_rollout(cuda, ns, name, precision, bar):
    """forward at the config's train shape + the 10-step rollout loop of utils/train_utils.py:203-207, eager
    fp32 reference vs accelerate()d copy of the same object, same GPU."""
    _, (ref_m, _, _), (fast_m, _, _) = _pair(ns, name, cuda, precision)
    ref_m.eval(), fast_m.eval()
    x, ib, _ = _batch(name, cuda)
    with torch.no_grad():
        y_ref, y = ref_m(x, ib), fast_m(x, ib)
        e_fwd = _rel(y, y_ref)
        seq_r = seq_f = x[:, :1]
        for i in range(10):     # the reference loop, line for line
            out_r = ref_m(seq_r, ib[:, : i + 1])
            seq_r = torch.cat((seq_r, out_r[:, -1:]), dim=1)
            out_f = fast_m(seq_f, ib[:, : i + 1])
            seq_f = torch.cat((seq_f, out_f[:, -1:]), dim=1)
        e_roll = _rel(seq_f[:, 1:], seq_r[:, 1:])
    print(f"\n[drop-in fwd] {name} {precision}: forward {tuple(x.shape)} rel {e_fwd:.2e}, 10-step rollout rel {e_roll:.2e}")
    assert e_fwd < bar and e_roll < bar


@pytest.mark.parametrize("name", list(CASES))
def test_bench_size_forward_against_the_reference_itself(cuda, ns, name):
    """BASELINE's full size (32 trajectories, prefix length 100 = the last and heaviest step of the benchmark's rollout),
    both configs at full width: the accelerated copy against the UNMODIFIED reference's eager fp32 forward on the same GPU —
    fp32 mode within 1e-4, bf16 mode within 2e-2, and the two modes of the library within 2e-2 of each other."""
    outs = {}
    g = torch.Generator(device="cpu").manual_seed(21)
    E = 1024 if name == "cylinder_flow" else 2048
    x = torch.randn(32, 100, 2, E, generator=g).to(cuda)
    ib = torch.rand(32, 1, 1, generator=g).expand(32, 100, 1).contiguous().to(cuda)
    def ms_of(fn):      # informational only (printed, never asserted): same GPU, same inputs, CUDA events
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3

    ms = {}
    for precision in ("fp32", "bf16"):
        _, (ref_m, _, _), (fast_m, _, _) = _pair(ns, name, cuda, precision)
        ref_m.eval(), fast_m.eval()
        with torch.no_grad():
            if "ref" not in outs:
                outs["ref"] = ref_m(x, ib)
                ms["ref"] = ms_of(lambda: ref_m(x, ib))
            outs[precision] = fast_m(x, ib)
            ms[precision] = ms_of(lambda: fast_m(x, ib))
        del ref_m, fast_m
        torch.cuda.empty_cache()
    e32, e16 = _rel(outs["fp32"], outs["ref"]), _rel(outs["bf16"], outs["ref"])
    print(f"\n[drop-in full size] {name} [32,100,2,{E}]: fp32 mode rel {e32:.2e}, bf16 mode rel {e16:.2e}; one forward: "
          f"reference eager fp32 on this GPU {ms['ref']:.2f} ms, accelerated fp32 mode {ms['fp32']:.2f} ms, bf16 mode {ms['bf16']:.2f} ms")
    assert e32 < 1e-4 and e16 < 2e-2 and _rel(outs["bf16"], outs["fp32"]) < 2e-2


GRAD_BARS = {"bf16": dict(loss=2e-3, dx=2e-2, grad=4e-2, cos=0.999), "fp32": dict(loss=1e-5, dx=1e-4, grad=2e-4, cos=0.999999)}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_accelerated_reference_loss_and_every_grad(cuda, ns, name, precision):
    """loss.backward() on the accelerated reference object: loss, dL/dx and EVERY parameter .grad against
    autograd through the eager reference at the config's full shape (tcgen05 attention backward at head dim
    128 / 256 self, 64 / 128 cross; K = 8192 / 16384 weight-gradient GEMMs)."""
    _, (ref_m, ref_loss, _), (fast_m, fast_loss, _) = _pair(ns, name, cuda, precision)
    ref_m.train(), fast_m.train()
    x, ib, tgt = _batch(name, cuda)
    xr = x.clone().requires_grad_(True)
    lr_ = ref_loss(ref_m(xr, ib), tgt)
    lr_.backward()
    xf = x.clone().requires_grad_(True)
    lf = fast_loss(fast_m(xf, ib), tgt)
    lf.backward()
    torch.cuda.synchronize()
    bars = GRAD_BARS[precision]
    e_loss = abs(lf.item() - lr_.item()) / abs(lr_.item())
    e_dx = _rel(xf.grad, xr.grad)
    ref_p, fast_p = dict(ref_m.named_parameters()), dict(fast_m.named_parameters())
    worst, worst_name, worst_cos, n_checked = 0.0, "", 1.0, 0
    for n, p in ref_p.items():
        if p.grad is None:
            assert fast_p[n].grad is None, f"dead parameter {n} got a gradient"   # 34 dead tensors
            continue
        g = fast_p[n].grad
        assert g is not None, n
        if p.grad.double().norm().item() < 1e-7:
            continue
        e = _rel(g, p.grad)
        c = F.cosine_similarity(g.flatten().double(), p.grad.flatten().double(), dim=0).item()
        n_checked += 1
        if e > worst:
            worst, worst_name = e, n
        worst_cos = min(worst_cos, c)
    print(f"\n[drop-in bwd] {name} {precision}: loss rel {e_loss:.2e}, dx rel {e_dx:.2e}, {n_checked} grads, "
          f"worst rel {worst:.2e} ({worst_name}), min cos {worst_cos:.6f}")
    assert e_loss < bars["loss"] and e_dx < bars["dx"]
    assert worst < bars["grad"] and worst_cos > bars["cos"]
    assert n_checked > 60


def _loaders(ns, name, n_traj, T, device, batch_size, seed=3):
    """DataLoader over the reference's own TemporalDataset (utils/data_processors.py:385-455) on synthetic encoded
    trajectories, in the layout get_datasets hands it (train/train_temporal.py:47-55): [tr, T+1, V, E] latents,
    originals, and the per-trajectory ib; same shuffle generator seed as the reference (:78-82)."""
    E = 1024 if name == "cylinder_flow" else 2048
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(n_traj, 1, 2, E, generator=g)
    # smooth synthetic dynamics so that there is something to learn: a damped rotation of a random state
    steps = [lat]
    for _ in range(T):
        steps.append(0.98 * torch.roll(steps[-1], 1, dims=-1) + 0.02 * torch.randn(n_traj, 1, 2, E, generator=g))
    lat = torch.cat(steps, dim=1)                                  # [tr, T+1, V, E]
    ib = torch.rand(n_traj, 1, 1, generator=g).expand(n_traj, T + 1, 1).contiguous()
    ds = ns.data_processors.TemporalDataset(lat, lat, ib, T, 0, str(device), False)
    gl = torch.Generator().manual_seed(42)
    return DataLoader(ds, batch_size=batch_size, shuffle=True, generator=gl)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("precision,bar", [("bf16", 1e-2), ("fp32", 1e-4)])
def test_reference_train_loop_100_iterations(cuda, ns, name, precision, bar):
    """100 iterations of the reference's loop body (train/train_temporal.py:252-260) — unchanged get_model through
    the hook, torch.optim.AdamW at the config's learning rate, config batch size and window — per-step loss of
    the CUDA path against the eager reference trained from the same init on the same batches."""
    cfg, (ref_m, ref_loss, ref_opt), (fast_m, fast_loss, fast_opt) = _pair(ns, name, cuda, precision)
    c = CASES[name]
    assert cfg["batch_size"] == c["B"] and cfg["dataset_src_len"] == c["T"]
    assert ref_opt.param_groups[0]["lr"] == cfg["learning_rate"] == fast_opt.param_groups[0]["lr"]
    losses = {}
    for tag, model, loss_fn, optimizer in (("ref", ref_m, ref_loss, ref_opt), ("fast", fast_m, fast_loss, fast_opt)):
        loader = _loaders(ns, name, 4 * c["B"], c["T"], cuda, cfg["batch_size"])
        device = cuda
        out, it = [], 0
        model.train()
        while it < 100:
            for data, target, _, ib in loader:      # ---- train/train_temporal.py:253-262, verbatim ----
                data, target, ib = data.to(device), target.to(device), ib.to(device)
                optimizer.zero_grad()
                outputs = model(data, ib)
                loss = loss_fn(outputs, target)
                loss.backward()
                optimizer.step()
                out.append(loss.item())
                it += 1
                if it == 100:
                    break
        losses[tag] = np.array(out)
    rel = np.abs(losses["fast"] - losses["ref"]) / np.abs(losses["ref"])
    print(f"\n[drop-in train] {name} {precision}: loss {losses['ref'][0]:.4f} -> {losses['ref'][-1]:.4f} (reference), "
          f"{losses['fast'][0]:.4f} -> {losses['fast'][-1]:.4f} (sea_b200); max per-step rel diff {rel.max():.2e} "
          f"(step {int(rel.argmax())}), mean {rel.mean():.2e}")
    assert losses["ref"][-1] < 0.8 * losses["ref"][0]
    assert rel.max() < bar
    # the weights the two optimizers ended on (same nn.Parameter objects the reference owns)
    pr, pf = dict(ref_m.named_parameters()), dict(fast_m.named_parameters())
    drift = max(_rel(pf[n], p) for n, p in pr.items() if p.requires_grad and p.numel() > 1024)
    print(f"    max relative weight difference after 100 AdamW steps: {drift:.2e}")
    assert drift < (5e-2 if precision == "bf16" else 1e-3)


@pytest.mark.parametrize("name", list(CASES))
def test_reference_rollout_loop_verbatim(cuda, ns, name):
    """utils/train_utils.py:154-185 ``autoregressive_validation`` called as is (model.eval(), no_grad, the
    prefix loop with torch.cat) on the eager and on the accelerated model: returned loss / relMSE agree."""
    for precision, bar in (("fp32", 1e-4), ("bf16", 3e-2)):
        cfg, (ref_m, ref_loss, _), (fast_m, fast_loss, _) = _pair(ns, name, cuda, precision)
        T = 24
        l_ref, r_ref = ns.train_utils.autoregressive_validation(ref_m, _loaders(ns, name, 2, T, cuda, 2), ref_loss, cuda)
        l_fast, r_fast = ns.train_utils.autoregressive_validation(fast_m, _loaders(ns, name, 2, T, cuda, 2), fast_loss, cuda)
        print(f"\n[drop-in rollout loop] {name} {precision}: loss {l_ref:.6f} vs {l_fast:.6f}, relMSE {r_ref:.6f} vs {r_fast:.6f}")
        assert abs(l_fast - l_ref) < bar * abs(l_ref) and abs(r_fast - r_ref) < bar * abs(r_ref)
        assert not fast_m.training


def test_train_mode_dropout_through_unchanged_loop(cuda, ns):
    """cylinder_flow ships dropout 0.1 (configs/cylinder_flow.py:120): the accelerated reference object trains
    through the unchanged loop body with the in-kernel masks (loss finite and decreasing, eval() is
    deterministic, train() is not)."""
    cfg, _, (m, loss_fn, opt) = _pair(ns, "cylinder_flow", cuda, "bf16", dropout=0.1)
    assert cfg["dropout"] == 0.1
    x, ib, tgt = _batch("cylinder_flow", cuda, T=64)
    m.train()
    losses = []
    for _ in range(30):
        opt.zero_grad()
        loss = loss_fn(m(x, ib), tgt)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    with torch.no_grad():
        a, b = m(x, ib), m(x, ib)
        assert not torch.equal(a, b)            # train mode under no_grad still drops (nn.Dropout semantics)
        m.eval()
        c, d = m(x, ib), m(x, ib)
        assert torch.equal(c, d)


def test_train_validate_train_validate_graphed_rollout(cuda, ns):
    """rollout() replaces the validation loop during training; the recorded CUDA graphs must
    stay valid (RoPE tables, packed weights) across train -> rollout -> train -> rollout."""
    from sea_b200.rollout import rollout
    _, (ref_m, _, _), (m, loss_fn, opt) = _pair(ns, "cylinder_flow", cuda, "bf16")
    x, ib, tgt = _batch("cylinder_flow", cuda, T=48)
    x0, steps = x[:, :1].contiguous(), 12

    def eager_loop(model):
        seq = x0
        with torch.no_grad():
            for i in range(steps):
                seq = torch.cat((seq, model(seq, ib[:, : i + 1])[:, -1:]), dim=1)
        return seq[:, 1:]

    for round_ in range(3):
        m.train()
        for _ in range(2):
            opt.zero_grad()
            loss_fn(m(x, ib), tgt).backward()
            opt.step()
        m.eval()
        got = rollout(m, x0, ib, steps)               # CUDA-graph plan (re-used across rounds when valid)
        want = rollout(m, x0, ib, steps, graphs=False)   # the same kernels launched eagerly
        assert torch.equal(got, want), f"round {round_}: graphed rollout differs from the eager launch sequence"
        want = eager_loop(m)                          # the unchanged loop through the rebound forward
        assert _rel(got, want) < 5e-3, f"round {round_}"
        got_c = rollout(m, x0, ib, steps, cached=True)
        assert _rel(got_c, want) < 2e-2
    ref_m.load_state_dict(m.state_dict())
    ref_m.eval()
    assert _rel(want, eager_loop(ref_m)) < 2e-2       # and it is still the reference's function of the new weights


def test_frozen_parameters_dx_only_and_double_backward(cuda, ns):
    """all parameters frozen -> dL/dx only; a second backward over the same graph raises."""
    _, (ref_m, _, _), (m, _, _) = _pair(ns, "cylinder_flow", cuda, "bf16")
    for p in list(m.parameters()) + list(ref_m.parameters()):
        p.requires_grad_(False)
    x, ib, tgt = _batch("cylinder_flow", cuda, T=32)
    xf = x.clone().requires_grad_(True)
    y = m(xf, ib)
    loss = F.mse_loss(y, tgt)
    loss.backward(retain_graph=True)
    xr = x.clone().requires_grad_(True)
    F.mse_loss(ref_m(xr, ib), tgt).backward()
    assert _rel(xf.grad, xr.grad) < 2e-2
    assert all(p.grad is None for p in m.parameters())
    with pytest.raises(RuntimeError, match="released"):
        loss.backward()


def test_external_update_repacks_weights(cuda, ns):
    """the fused AdamW constructed WITHOUT engine= bumps parameter versions, so the engine
    re-packs its bf16 copies instead of silently running on stale ones."""
    from sea_b200.optim import AdamW
    _, (ref_m, _, _), (m, loss_fn, _) = _pair(ns, "multiphase_flow", cuda, "bf16")
    opt = AdamW(m.parameters(), lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)   # no engine=
    x, ib, tgt = _batch("multiphase_flow", cuda, T=32)
    m.train()
    y0 = m(x, ib).detach().clone()
    opt.zero_grad()
    loss_fn(m(x, ib), tgt).backward()
    opt.step()
    y1 = m(x, ib).detach()
    assert _rel(y1, y0) > 1e-3                        # the update is visible in the next forward
    ref_m.load_state_dict(m.state_dict())
    ref_m.train()
    with torch.no_grad():
        assert _rel(y1, ref_m(x, ib)) < 2e-2          # ... and it is the forward of the UPDATED masters


@pytest.mark.parametrize("precision,bar", [("fp32", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("name", list(CASES))
def test_accelerated_reference_spatial_model(cuda, ns, name, precision, bar):
    """accelerate_spatial on the reference's own SpatialModel built by ProcessData.initialize_spatial_model
    (utils/data_processors.py:305-317) through the hook: encode / decode / forward vs the eager copy."""
    import sea_b200
    cfg = _cfg(name, cuda)
    torch.manual_seed(1)
    proc = ns.data_processors.ProcessData(64, cfg)
    eager = proc.initialize_spatial_model()
    sea_b200.install(spatial=True, spatial_precision=precision)
    try:
        fast = proc.initialize_spatial_model()
    finally:
        sea_b200.uninstall()
    assert type(fast) is ns.encoder_decoder.SpatialModel and hasattr(fast, "_sea_codec")
    fast.load_state_dict(eager.state_dict())
    eager.eval(), fast.eval()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(37, 64, 3, 64, generator=g)
    x[:, :, :, 50:] = 0.0
    x[0, 0, 0, -1] = -9999.0
    xa, xb = x.to(cuda), x.to(cuda)
    with torch.no_grad():
        ya, yb = eager(xa), fast(xb)
        assert torch.equal(xa, xb)                         # generate_padding_mask rewrote both in place
        za, zb = eager.encode(xa), fast.encode(xb)
        da, db = eager.decode(za), fast.decode(za)
    print(f"\n[drop-in spatial] {name} {precision}: forward rel {_rel(yb, ya):.2e}, encode {_rel(zb, za):.2e}, "
          f"decode {_rel(db, da):.2e}")
    assert _rel(yb, ya) < bar and _rel(zb, za) < bar and _rel(db, da) < bar


def test_eager_loop_condition_detection(cuda, ns):
    """The unchanged loop model(seq, ib[:, :i+1]) through the drop-in forward: the engine tests the BASE condition tensor once
    for time-invariance and then takes the per-trajectory condition path with cached condition rows.  It must (i) agree with
    the eager reference for a time-invariant ib, (ii) NOT take that path for a time-varying ib, (iii) notice an in-place
    update of ib (version counter) and a different batch slice of the same base tensor."""
    _, (ref_m, _, _), (fast_m, _, _) = _pair(ns, "cylinder_flow", cuda, "fp32")
    ref_m.eval(), fast_m.eval()
    eng = fast_m._sea_engine
    g = torch.Generator(device="cpu").manual_seed(17)
    B, steps = 6, 7
    x0 = torch.randn(B, 1, 2, 1024, generator=g).to(cuda)

    def loop(m, x0_, ib_):
        seq = x0_
        with torch.no_grad():
            for i in range(steps):     # utils/train_utils.py:203-207
                out = m(seq, ib_[:, : i + 1])
                seq = torch.cat((seq, out[:, -1:]), dim=1)
        return seq[:, 1:]

    ib_inv = torch.rand(B, 1, 1, generator=g).expand(B, steps, 1).contiguous().to(cuda)
    assert _rel(loop(fast_m, x0, ib_inv), loop(ref_m, x0, ib_inv)) < 1e-4
    assert eng._ib_auto is True
    ib_inv.mul_(0.5)                                       # in place: same storage, new version
    assert _rel(loop(fast_m, x0, ib_inv), loop(ref_m, x0, ib_inv)) < 1e-4
    # a batch slice of the same base tensor: other trajectories, other condition rows
    assert _rel(loop(fast_m, x0[2:5], ib_inv[2:5]), loop(ref_m, x0[2:5], ib_inv[2:5])) < 1e-4
    assert _rel(loop(fast_m, x0[:3], ib_inv[:3]), loop(ref_m, x0[:3], ib_inv[:3])) < 1e-4
    ib_var = torch.rand(B, steps, 1, generator=g).to(cuda)
    assert _rel(loop(fast_m, x0, ib_var), loop(ref_m, x0, ib_var)) < 1e-4
    assert eng._ib_auto is False
    # a NEW tensor that the caching allocator places at the SAME address (next batch of a validation loop), other values:
    # the cache is tied to the tensor object, never to its address
    addr = None
    for k in range(4):
        vals = torch.rand(B, 1, 1, generator=g)
        ib_new = (vals.expand(B, steps, 1) if k % 2 == 0 else torch.rand(B, steps, 1, generator=g)).contiguous().to(cuda)
        if addr is not None:
            assert ib_new.data_ptr() == addr            # same address as the tensor freed in the previous iteration
        addr = ib_new.data_ptr()
        assert _rel(loop(fast_m, x0, ib_new), loop(ref_m, x0, ib_new)) < 1e-4, k
        assert eng._ib_auto is (k % 2 == 0)
        del ib_new
