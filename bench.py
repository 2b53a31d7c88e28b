#!/usr/bin/env python
"""bench.py — headline benchmark of the SEA hot path on B200 (see BASELINE.json).

Metric: autoregressive rollout throughput of the cylinder_flow temporal model, in
trajectory-steps per second (one unit = one model step of one trajectory).  A bench "step" is one
complete R-step rollout of the per-GPU ensemble of B trajectories through the module's public
``forward`` with the reference's loop semantics (utils/train_utils.py:202-209: whole prefix
recomputed every step, no KV cache).  Trajectories are sharded across GPUs (weak scaling: B per
GPU fixed), no data-path collective.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
  python bench.py --impl reference [...]                         # CPU arm: oracle port, host cores

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(name="cylinder_flow", num_layers=1, embed_dim=1024, n_heads=8, max_len=2024, scale_ratio=8,
           src_len=0, num_variables=2, down_proj=2, ln_type="adaln")
B_PER_GPU = 32          # trajectories per GPU
ROLLOUT_STEPS = 100     # autoregressive steps per trajectory
CPU_SAMPLE = dict(B=8, R=40)   # bounded sample of the same workload for the CPU arm
WORKLOAD = (f"cylinder_flow temporal model (configs/cylinder_flow.py: E=1024, 8 heads, H=8192, Dd=512, "
            f"V=2, adaln), {ROLLOUT_STEPS}-step autoregressive rollout with full-prefix recompute, "
            f"{B_PER_GPU} trajectories per GPU")
METRIC = "rollout_trajectory_steps_per_sec"
UNIT = "trajectory-steps/s"


def fwd_flops(B, T, E=1024, H=8192, Dd=512, V=2, L=1, adaln=True, c=0.5):
    """SURVEY.md §8(d) formula (c = 1/2: causal-useful attention FLOPs)."""
    M = B * T
    per_layer = (8 * M * E * E + c * 4 * B * T * T * E
                 + (V - 1) * (4 * M * E * Dd + 8 * M * Dd * Dd + c * 4 * B * T * T * Dd + 2 * M * Dd * E)
                 + 4 * M * E * H + 2 * M * E * E + 16 * M + 16 * M * E)
    if adaln:
        per_layer += 2 * (8 * M * E * E + 4 * M * E) + (V - 1) * 2 * (8 * M * Dd * Dd + 4 * M * Dd)
    final = (8 * M * E * E + 4 * M * E) if adaln else 0
    return V * (L * per_layer + final)


def rollout_flops(B, R):
    return sum(fwd_flops(B, t) for t in range(1, R + 1))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], src="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


def gemm_traffic():
    """DRAM bytes per GEMM launch from the committed `ncu --set full` capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r1c_gemm_traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def make_inputs(B, R, E, V, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, 1, V, E, generator=g)
    ib = torch.rand(B, 1, 1, generator=g).expand(B, R, 1).contiguous()  # time-invariant parameter
    return x0, ib


def build_model(precision="bf16"):
    from sea_b200.temporal import TemporalModel
    torch.manual_seed(42)
    c = CFG
    return TemporalModel(c["num_layers"], c["embed_dim"], c["n_heads"], c["max_len"], c["scale_ratio"],
                         c["src_len"], c["num_variables"], c["down_proj"], 0.0, "sea", "learnable", "mlp",
                         "add", 1, 1, True, c["ln_type"], precision=precision)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # median over the busier half of the samples (the sampler also sees the idle edges)
        busy = sm[len(sm) // 2:] if sm else []
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def cpu_rollout_rate(sd, B, R, reps, warm):
    """Oracle port of the reference path on the host cores: trajectory-steps/s."""
    from oracle import sea_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    x0, ib = make_inputs(B, R, CFG["embed_dim"], CFG["num_variables"], 1234)
    kw = dict(num_layers=1, n_heads=CFG["n_heads"], ln_type=CFG["ln_type"])
    times = []
    with torch.no_grad():
        for i in range(warm + reps):
            t0 = time.perf_counter()
            so.rollout(x0, ib, R, sd, **kw)
            if i >= warm:
                times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return B * R / med, med, torch.get_num_threads()


def train_step_aux(dev, timed):
    """multiphase_flow (E=2048, hd=256, H=16384, ln), B=4, T=199 (configs/multiphase_flow.py:140-141):
    zero_grad + forward + MSE + backward + fused AdamW, device-resident synthetic batch."""
    from sea_b200.optim import AdamW
    from sea_b200.temporal import TemporalModel
    torch.manual_seed(42)
    m = TemporalModel(1, 2048, 8, 2024, 8, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, "ln").to(dev).train()
    opt = AdamW(m.parameters(), lr=8e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, engine=m.engine())
    B, T = 4, 199
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(B, T, 2, 2048, device=dev, generator=g)
    ib = torch.rand(B, 1, 1, device=dev, generator=g).expand(B, T, 1).contiguous()
    tgt = torch.randn(B, T, 2, 2048, device=dev, generator=g)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(m(x, ib), tgt)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    ms = timed(step, 5)
    flops = 3 * fwd_flops(B, T, E=2048, H=16384, Dd=1024, adaln=False)
    out = {"workload": "multiphase_flow train step (B=4, T=199): zero_grad + fwd + MSE + bwd + fused AdamW, bf16",
           "ms_per_step": ms, "samples_per_sec": B / (ms / 1e3), "model_tflops": flops / (ms * 1e-3) / 1e12}
    del m, opt
    torch.cuda.empty_cache()
    return out


def attention_kernel_aux(dev, timed, pk):
    """BASELINE metric, second half ("attention TFLOP/s vs peak"): the fused causal attention kernels
    alone at the configs' max_len (T = 2024), cylinder_flow head geometry (8 heads x 128), bf16,
    causal-useful FLOPs 2*B*nh*T^2*hd forward (c = 1/2), 2.5x that backward."""
    from sea_b200 import ops
    B, T, nh, hd = 4, 2024, 8, 128
    g = torch.Generator(device=dev).manual_seed(9)
    qkv = torch.randn(B * T, 3 * nh * hd, device=dev, generator=g).bfloat16()
    q, k, v = qkv[:, : nh * hd], qkv[:, nh * hd: 2 * nh * hd], qkv[:, 2 * nh * hd:]
    o, lse = ops.attention_fwd(q, k, v, nh, B=B, want_lse=True)
    do = torch.randn(B * T, nh * hd, device=dev, generator=g).bfloat16()
    for _ in range(3):
        ops.attention_fwd(q, k, v, nh, B=B, want_lse=True)
        ops.attention_bwd(q, k, v, o, do, lse, nh, B=B)
    ms_f = timed(lambda: ops.attention_fwd(q, k, v, nh, B=B, want_lse=True), 20)
    ms_b = timed(lambda: ops.attention_bwd(q, k, v, o, do, lse, nh, B=B), 10)
    fl = 2.0 * B * nh * T * T * hd
    tf_f, tf_b = fl / (ms_f * 1e-3) / 1e12, 2.5 * fl / (ms_b * 1e-3) / 1e12
    return {"workload": f"causal self-attention B={B} T={T} heads={nh} head_dim={hd}, bf16 (tcgen05 kernels alone)",
            "fwd_ms": ms_f, "fwd_tflops": tf_f, "fwd_frac_of_sustained_peak": tf_f / pk["sustained"],
            "bwd_ms": ms_b, "bwd_tflops": tf_b, "bwd_frac_of_sustained_peak": tf_b / pk["sustained"],
            "flops": "causal-useful (half of the dense count)"}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path.  /root/reference is Python
    and does not travel to the GPU box, so this times oracle/sea_oracle.py (the restatement pinned
    against the reference's outputs) on all host cores, on a bounded sample of the workload."""
    if rank != 0:
        return
    torch.manual_seed(42)
    sd = {k: v for k, v in build_model().state_dict().items()}
    B, R = CPU_SAMPLE["B"], CPU_SAMPLE["R"]
    from oracle import sea_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    x0, ib = make_inputs(B, R, CFG["embed_dim"], CFG["num_variables"], 1234)
    kw = dict(num_layers=1, n_heads=CFG["n_heads"], ln_type=CFG["ln_type"])
    with torch.no_grad():
        for _ in range(args.warmup):
            so.rollout(x0, ib, R, sd, **kw)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            so.rollout(x0, ib, R, sd, **kw)
        dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    val = B * R / (ms / 1e3)
    sample = (f"{B} trajectories x {R} steps of the same rollout (prefix recompute) per bench step, fp32, "
              f"torch CPU ({torch.get_num_threads()} threads)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--rollout", type=int, default=ROLLOUT_STEPS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from sea_b200 import lib
    from sea_b200.rollout import profile, rollout, rollout_from_host

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, R = args.batch, args.rollout
    E, V = CFG["embed_dim"], CFG["num_variables"]
    model = build_model(args.precision)
    sd_cpu = {k: v.clone() for k, v in model.state_dict().items()} if rank == 0 else None
    model = model.to(dev).eval()
    eng = model.engine()

    x0_h, ib_h = make_inputs(B, R, E, V, 1234 + rank)          # each rank owns its trajectories
    x0_h, ib_h = x0_h.pin_memory(), ib_h.pin_memory()
    out_h = torch.empty(B, R, V, E).pin_memory()
    x0, ib = x0_h.to(dev), ib_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    # ---- device-resident leg -------------------------------------------------------------
    for _ in range(args.warmup):
        rollout(model, x0, ib, R)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.total_launches
    ms_step = timed(lambda: rollout(model, x0, ib, R), args.steps)
    gpu_launches = eng.total_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * R / (ms_step / 1e3)

    # ---- end-to-end leg: pinned host buffers in / out, copies inside the timed region -------
    for _ in range(2):
        rollout_from_host(model, x0_h, ib_h, R, out_h, dev)
    ms_e2e = timed(lambda: rollout_from_host(model, x0_h, ib_h, R, out_h, dev), args.steps)
    e2e_value = world * B * R / (ms_e2e / 1e3)
    h2d = x0_h.numel() * 4 + ib_h.numel() * 4
    d2h = out_h.numel() * 4

    # ---- KV-cached incremental engine (opt-in, NOT the headline): same trajectories ----
    pred_prefix = rollout(model, x0, ib, R)
    for _ in range(2):
        pred_cached = rollout(model, x0, ib, R, cached=True)
    cached_rel = ((pred_cached - pred_prefix).norm() / pred_prefix.norm()).item()
    n10 = min(10, R)
    cached_rel10 = ((pred_cached[:, :n10] - pred_prefix[:, :n10]).norm() / pred_prefix[:, :n10].norm()).item()
    ms_cached = timed(lambda: rollout(model, x0, ib, R, cached=True, _view_ok=True), args.steps)
    del pred_prefix, pred_cached

    # ---- auxiliary: BASELINE configs[1], multiphase_flow fwd+bwd(+AdamW) train step at the reference batch ----
    train_aux = train_step_aux(dev, timed) if world == 1 else None   # single-GPU figure only

    # ---- roofline leg: one more rollout with per-launch CUDA events on the launch stream ----
    pk = peaks()
    attn_aux = attention_kernel_aux(dev, timed, pk)
    with profile() as prof:
        rollout(model, x0, ib, R)
        torch.cuda.synchronize()
    ps = prof.summary
    gemm_tflops = ps["gemm"]["work"] / (ps["gemm"]["ms"] * 1e-3) / 1e12 if ps["gemm"]["ms"] > 0 else 0.0
    attn_tflops = ps["attention"]["work"] / (ps["attention"]["ms"] * 1e-3) / 1e12 if ps["attention"]["ms"] > 0 else 0.0
    elem_gbs = ps["elementwise"]["work"] / (ps["elementwise"]["ms"] * 1e-3) / 1e9 if ps["elementwise"]["ms"] > 0 else 0.0
    total_kernel_ms = sum(v["ms"] for v in ps.values())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, med, cores = cpu_rollout_rate(sd_cpu, CPU_SAMPLE["B"], CPU_SAMPLE["R"], reps=2, warm=1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": (f"oracle port (torch CPU fp32) of the same rollout on {CPU_SAMPLE['B']} trajectories x "
                          f"{CPU_SAMPLE['R']} steps, median of 2 after 1 warm-up ({med:.2f} s each)")}

    flops = rollout_flops(B, R)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision if args.precision != "fp32" else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD if (B, R) == (B_PER_GPU, ROLLOUT_STEPS) else
                   WORKLOAD + f" [override: B={B}, R={R}]",
                   "l2": "no flush: per-step working set (174 MB bf16 weights + activations up to "
                         ">1 GB) exceeds the 126 MB L2",
                   "bench_step": f"one {R}-step rollout of {B} trajectories per GPU",
                   "execution": "sea_b200.rollout.RolloutPlan: one CUDA graph per prefix length (full forward "
                                "over the prefix read in place from the sequence buffer, append last step); the roofline leg replays the "
                                "same kernels eagerly with per-launch CUDA events",
                   "algorithmic_tflop_per_step_per_gpu": flops / 1e12,
                   "model_tflops_per_gpu": flops / (ms_step * 1e-3) / 1e12},
        "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": pk["sustained"],
                     "unit": "TFLOP/s", "frac": gemm_tflops / pk["sustained"], "traffic": gemm_traffic(),
                     "traffic_note": "DRAM bytes per GEMM launch of the T=100 forward, ncu --set full "
                                     "(profiles/r1c_summary.md section 2); not measured in this run",
                     "kernel": "gemm_bf16_tn_kernel (tcgen05)", "peak_source": pk["src"] + ", sustained",
                     "gemm_share_of_kernel_time": ps["gemm"]["ms"] / total_kernel_ms if total_kernel_ms else None,
                     "attention_tflops": attn_tflops, "attention_frac": attn_tflops / pk["sustained"],
                     "elementwise_gbs": elem_gbs, "elementwise_frac_of_hbm": elem_gbs / pk["hbm"],
                     "breakdown_ms": {k: v["ms"] for k, v in ps.items()},
                     "launches": {k: v["launches"] for k, v in ps.items()}},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "train_step": train_aux,
        "attention_kernel": attn_aux,
        "cached_rollout": {"value": world * B * R / (ms_cached / 1e3), "unit": UNIT, "ms_per_step": ms_cached,
                           "rel_l2_vs_prefix_loop": cached_rel, "rel_l2_vs_prefix_loop_first_10_steps": cached_rel10,
                           "note": "opt-in KV-cached engine (sea_temporal_step): O(1) work per step instead of "
                                   "the reference loop's prefix recompute; same outputs up to rounding (bf16 rounding "
                                   "differences grow along a 100-step autoregressive rollout; in fp32 mode the two "
                                   "engines agree to 2e-7, tests/test_temporal_gpu.py); not the headline value"},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
