"""Data-parallel train-step timing (BASELINE configs[2]: cylinder_flow temporal training, bf16, DP over N GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 scripts/dp_train_bench.py [--config multiphase_flow] [--b 2]

Per rank: zero_grad + forward + MSE + backward + gradient exchange + fused AdamW on its own b trajectories
(weak scaling).  Reports, max over ranks, CUDA events: the step without any exchange, with the exchange after the
backward, and with the stream-MLP bucket overlapped with the rest of the backward (sea_b200.parallel.train_step).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from sea_b200 import parallel  # noqa: E402
from sea_b200.optim import AdamW  # noqa: E402
from sea_b200.temporal import TemporalModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cylinder_flow")
ap.add_argument("--b", type=int, default=0)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
rank, world, local = parallel.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if args.config == "cylinder_flow":
    E, scale, ln, T, b, lr = 1024, 8, "adaln", 399, args.b or 2, 1e-4
else:
    E, scale, ln, T, b, lr = 2048, 8, "ln", 199, args.b or 4, 8e-5
torch.manual_seed(42)
m = TemporalModel(1, E, 8, 2024, scale, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln).to(dev).train()
opt = AdamW(m.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, engine=m.engine())
g = torch.Generator(device=dev).manual_seed(1234 + rank)
x = torch.randn(b, T, 2, E, device=dev, generator=g)
ib = torch.rand(b, 1, 1, device=dev, generator=g).expand(b, T, 1).contiguous()
tgt = torch.randn(b, T, 2, E, device=dev, generator=g)


def timed(fn, steps):
    for _ in range(3):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def local_step():
    opt.zero_grad(set_to_none=True)
    F.mse_loss(m(x, ib), tgt).backward()
    opt.step()


ms_local = timed(local_step, args.steps)
ms_seq = timed(lambda: parallel.train_step(m, opt, F.mse_loss, x, tgt, ib, overlap=False), args.steps)
ms_ovl = timed(lambda: parallel.train_step(m, opt, F.mse_loss, x, tgt, ib, overlap=True), args.steps)
ms_ovl_f32 = timed(lambda: parallel.train_step(m, opt, F.mse_loss, x, tgt, ib, overlap=True, grad_dtype="f32"), args.steps)
eng = m.engine()
flat = eng.flat_grad()
if rank == 0:
    print(json.dumps({"config": args.config, "n_gpus": world, "per_gpu_batch": b, "T": T,
                      "grad_bytes": flat.numel() * 4, **parallel.TrainStep(m, opt, F.mse_loss).info(),
                      "ms_step_no_exchange": ms_local, "ms_step_exchange_after_backward": ms_seq,
                      "ms_step_exchange_overlapped": ms_ovl, "ms_step_exchange_overlapped_f32_buckets": ms_ovl_f32,
                      "samples_per_sec_overlapped": world * b / (ms_ovl / 1e3)}))
if world > 1:
    dist.destroy_process_group()
