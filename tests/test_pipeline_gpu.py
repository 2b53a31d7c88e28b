"""Resident fields -> latents -> rollout -> fields pipeline (sea_b200.pipeline) against the reference's OWN
chain run beside it: MeshProcessor.patchify_and_scale (utils/data_processors.py:484-542) -> SpatialModel.encode ->
transform_processed_data (utils/train_utils.py:315-338) -> the rollout loop (:202-209) -> inverse_transform_processed_data
(:340-362) -> SpatialModel.decode -> MeshProcessor.inverse_scale_and_unpatch (:553-573), on the configs' own model sizes."""
import numpy as np
import pytest
import torch

from oracle import ref as oref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not oref.available(), reason="reference not staged (oracle/_ref)")]


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ns():
    return oref.load()


def _mesh(n_cells, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n_cells, generator=g) * 2.2 - 0.3
    y = torch.rand(n_cells, generator=g) * 0.41
    return x, y


@pytest.mark.parametrize("name,precision,bar", [("cylinder_flow", "fp32", 2e-4), ("cylinder_flow", "bf16", 2e-2),
                                                ("multiphase_flow", "fp32", 2e-4), ("multiphase_flow", "bf16", 2e-2)])
def test_fields_to_fields_rollout_matches_reference_chain(cuda, ns, tmp_path, name, precision, bar):
    from sea_b200.pipeline import ResidentPipeline
    from sea_b200.spatial import accelerate_spatial
    from sea_b200.temporal import accelerate
    cfg = oref.temporal_config(name)
    cfg.update(device=str(cuda), dropout=0.0, save_dir=str(tmp_path), perform_initial_test=False)
    tr, T, steps, n_cells, F = 3, 7, 6, 2500, 3
    xc, yc = _mesh(n_cells, 3)
    g = torch.Generator().manual_seed(11)
    fields = torch.randn(tr, T, n_cells, F, generator=g) * torch.tensor([1.0, 0.3, 2.0]) + torch.tensor([0.5, 0.0, -1.0])
    ib = torch.rand(tr, 1, 1, generator=g).expand(tr, T, 1).contiguous()

    # ---------------- the reference chain (mesh work on the CPU, models on the GPU, as train_temporal.py runs it)
    mp = ns.data_processors.MeshProcessor(cfg, torch.stack((xc, yc)))
    _, scaled = mp.patchify_and_scale(fields.reshape(tr * T, n_cells, F), train_indices=np.arange(tr))   # [S,P,C,F]
    n_inp, P = scaled.shape[2], scaled.shape[1]
    torch.manual_seed(1)
    proc = ns.data_processors.ProcessData(n_inp, cfg)
    sp_ref = proc.initialize_spatial_model().eval()
    sp_fast = proc.initialize_spatial_model().eval()
    sp_fast.load_state_dict(sp_ref.state_dict())
    accelerate_spatial(sp_fast)
    torch.manual_seed(2)
    tm_ref, _, _ = ns.train_temporal.get_model(cfg, cuda)
    tm_fast, _, _ = ns.train_temporal.get_model(cfg, cuda)
    tm_fast.load_state_dict(tm_ref.state_dict())
    accelerate(tm_fast, precision=precision)
    tm_ref.eval(), tm_fast.eval()
    G = len(cfg["field_groups"])
    with torch.no_grad():
        x_in = scaled.permute(0, 1, 3, 2).contiguous().to(cuda)                       # SEA_isolate, train_temporal.py:153-154
        z = sp_ref.encode(sp_ref.generate_padding_mask(x_in))
        lat_ref = ns.train_utils.transform_processed_data(z, tr, T, P, G)            # [tr,T,G,P*D]
        seq = lat_ref[:, 0:1]
        ibd = ib.to(cuda)
        for i in range(steps):                                                      # utils/train_utils.py:203-207
            out = tm_ref(seq, ibd[:, :i + 1])
            seq = torch.cat((seq, out[:, -1:]), dim=1)
        roll_ref = seq[:, 1:]
        dec = ns.train_utils.inverse_transform_processed_data(roll_ref, tr, steps, P, G)
        dec = sp_ref.decode(dec).permute(0, 1, 3, 2)                                  # :225-226 -> [S,P,C,F]
        rec_ref = mp.inverse_scale_and_unpatch(dec.cpu()).reshape(tr, steps, n_cells, F)

    # ---------------- the resident pipeline
    pipe = ResidentPipeline(tm_fast, sp_fast, xc, yc, cfg["field_groups"], feature_range=cfg.get("scale_feature_range"),
                            m=cfg["m"], n=cfg["n"], device=cuda)
    fd = fields.to(cuda)
    patches = pipe.patchify(fd.reshape(tr * T, n_cells, F))
    assert torch.equal(patches.cpu(), scaled.permute(0, 1, 3, 2))                   # index work: bit-exact
    lat = pipe.encode_fields(fd)
    assert lat.shape == lat_ref.shape and _rel(lat, lat_ref) < 1e-4
    rec, roll = pipe.rollout_fields(fd[:, :1], ibd, steps, return_latents=True)
    e_lat, e_rec = _rel(roll, roll_ref), _rel(rec.cpu(), rec_ref)
    print(f"\n[pipeline] {name} {precision}: latents after {steps} steps rel {e_lat:.2e}, decoded fields rel {e_rec:.2e}")
    assert rec.shape == (tr, steps, n_cells, F)
    assert e_lat < bar and e_rec < bar
    # decode + unpatch alone, from the reference's latents: the codec's 1e-4 bar
    rec2 = pipe.decode_latents(roll_ref)
    assert _rel(rec2.cpu(), rec_ref) < 1e-4
    # the KV-cached engine through the same pipeline
    rec_c = pipe.rollout_fields(fd[:, :1], ibd, steps, cached=True)
    assert _rel(rec_c.cpu(), rec_ref) < bar


def test_scaled_gather_scatter_bit_exact_with_reference_minmax_scaler(cuda, ns, tmp_path):
    """MinMaxScaler.transform / inverse_transform (utils/data_processors.py:245-272) fused into the patch gather / scatter:
    same fp32 operation sequence, so the results are bit-identical to torch's; the round trip restores every cell."""
    from sea_b200.pipeline import ResidentPipeline
    from sea_b200.spatial import SpatialModel
    S_, n_cells, F = 5, 1777, 3
    xc, yc = _mesh(n_cells, 9)
    g = torch.Generator().manual_seed(4)
    fields = torch.randn(S_, n_cells, F, generator=g) * torch.tensor([3.0, 0.01, 40.0]) + torch.tensor([1.0, -0.2, 7.0])
    groups, fr = [[0, 1], [2]], (-1, 1)
    scalers = [ns.data_processors.MinMaxScaler(feature_range=fr, name=f"g{i}", save_dir=str(tmp_path)) for i in range(2)]
    scaled = torch.zeros_like(fields)
    for sc, grp in zip(scalers, groups):
        sc.fit(fields[:, :, grp])
        scaled[..., grp] = sc.transform(fields[..., grp])
    part = ns.data_processors.DataPartitioner2D(xc, yc, m=9, n=9, pad_id=-1, pad_field_value=0)
    pp, _ = part.create_partitions([scaled[:, :, i] for i in range(F)])
    want = torch.stack([p[1] for p in pp], dim=1)                                    # [S,P,C,F]
    cap = want.shape[2]
    sp = SpatialModel(groups, cap, 48, 1, 8, 8, 64, 0).to(cuda)
    pipe = ResidentPipeline(None, sp, xc, yc, groups, feature_range=fr, device=cuda)
    pipe.set_scalers([(s.min_val.item(), s.max_val.item()) for s in scalers])
    got = pipe.patchify(fields.to(cuda))
    assert torch.equal(got.cpu(), want.permute(0, 1, 3, 2))
    pipe2 = ResidentPipeline(None, sp, xc, yc, groups, feature_range=fr, device=cuda)
    pipe2.fit_scalers(fields.to(cuda))
    assert pipe2.minmax == pipe.minmax
    # inverse: the reference unpatches then un-scales
    _, rec = part.inverse_partition(pp, time_dim=S_)
    back = torch.zeros_like(rec)
    for sc, grp in zip(scalers, groups):
        back[..., grp] = sc.inverse_transform(rec[..., grp])
    got_back = pipe.unpatchify(got)
    assert torch.equal(got_back.cpu(), back)
    assert _rel(got_back.cpu(), fields) < 1e-6
