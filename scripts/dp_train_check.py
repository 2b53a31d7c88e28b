"""Data-parallel training check (run under torchrun, one rank per GPU):
every rank trains the SAME small temporal model on ITS shard of a global batch with the CUDA path
(sea_b200.parallel.train_step: backward -> all-reduce of the gradient buckets (NCCL; gloo with SEA_DP_BACKEND=gloo,
which also lets two ranks share one GPU) -> AdamW);
rank 0 also trains the fp32 oracle on CPU on the WHOLE global batch and compares the loss curves.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/dp_train_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import golden_recipe as gr  # noqa: E402
from oracle import sea_oracle as so  # noqa: E402
from sea_b200 import parallel  # noqa: E402
from sea_b200.temporal import TemporalModel  # noqa: E402

# SEA_DP_BACKEND=gloo: the ranks may SHARE a GPU (NCCL refuses two ranks on one device) — the CUDA backward with its
# per-group events, the bucket planner, the bf16 twin and the twin-reading AdamW are exactly those of the NCCL run; only
# the transport of the collective differs (gloo stages the CUDA buckets through pinned host memory)
rank, world, local = parallel.init_from_env(os.environ.get("SEA_DP_BACKEND") or None)
dev = torch.device("cuda", local % torch.cuda.device_count())
torch.cuda.set_device(dev)
E, nh, scale, V, T, steps = 128, 2, 2, 2, 12, 30
LR = 1e-3
if os.environ.get("SEA_DP_WIDTH", "small") == "full":    # cylinder_flow / multiphase_flow widths (short T: the oracle runs on CPU)
    LR = 1e-4 if os.environ.get("SEA_LN", "adaln") == "adaln" else 8e-5        # configs/*.py:147
    E, nh, scale, T, steps = (1024 if os.environ.get("SEA_LN", "adaln") == "adaln" else 2048), 8, 8, 16, 12
FUSED = os.environ.get("SEA_DP_OPT", "torch") == "fused"   # fused AdamW stepping from the averaged bf16 bucket
Bg = 2 * world                                            # global batch, 2 trajectories per rank
ln = os.environ.get("SEA_LN", "adaln")
OVERLAP = os.environ.get("SEA_DP_OVERLAP", "1") == "1"   # force the overlapped exchange (auto would skip it at this size)
sd = gr.fill_state(gr.temporal_shapes(embed_dim=E, n_heads=nh, scale_ratio=scale, num_variables=V, ln_type=ln), 11)
x, ib, tgt = gr.temporal_inputs(Bg, T, V, E, 11)
m = TemporalModel(1, E, nh, 64, scale, 0, V, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln)
m.load_state_dict(sd, strict=False)
m = m.to(dev).train()
if FUSED:
    from sea_b200.optim import AdamW
    opt = AdamW(m.parameters(), lr=LR, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, engine=m.engine())
else:
    opt = torch.optim.AdamW(m.parameters(), lr=LR, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
xs, ibs, ts = (parallel.shard_trajectories(t, rank, world).to(dev) for t in (x, ib, tgt))
losses = []
for _ in range(steps):
    loss = parallel.train_step(m, opt, F.mse_loss, xs, ts, ibs, overlap=OVERLAP)
    lt = loss.detach().clone()
    if world > 1:
        dist.all_reduce(lt)
    losses.append(lt.item() / world)                     # mean over equal shards == global MSE
# all ranks must hold identical weights after training
# (dead parameters of the reference never train and are randomly initialised per process: compare live ones)
w = torch.cat([p.detach().flatten() for p in m.parameters() if p.grad is not None])
ref = w.clone()
if world > 1:
    dist.broadcast(ref, 0)
drift = (w - ref).abs().max().item()
ok = True
if rank == 0:
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o = torch.optim.AdamW(list(leaves.values()), lr=LR, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    ref_losses = []
    for _ in range(steps):
        o.zero_grad()
        l = F.mse_loss(so.temporal_forward(x, ib, leaves, num_layers=1, n_heads=nh, ln_type=ln), tgt)
        l.backward()
        o.step()
        ref_losses.append(l.item())
    rel = max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses))
    print(f"DP_TRAIN world={world} loss {ref_losses[0]:.4f}->{ref_losses[-1]:.4f} (oracle, global batch {Bg}) "
          f"{losses[0]:.4f}->{losses[-1]:.4f} (cuda, {world} ranks); max per-step rel diff {rel:.2e}")
    ok = rel < 2e-2
print(f"rank {rank}: weight drift vs rank 0 = {drift:.3e}", flush=True)
flag = torch.tensor([1.0 if (ok and drift == 0.0) else 0.0], device=dev)
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
if rank == 0:
    print("DP_TRAIN_OK" if flag.item() == 1.0 else f"DP_TRAIN_FAIL drift={drift}")
sys.exit(0 if flag.item() == 1.0 else 1)
