"""N>1 on real GPUs (skipped on a 1-GPU box): data-parallel training parity and sharded rollout bench."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, script_args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)


@pytest.mark.parametrize("width,opt", [("small", "torch"), ("small", "fused"), ("full", "fused")])
@pytest.mark.parametrize("ln", ["adaln", "ln"])
def test_dp_training_matches_oracle_global_batch(cuda, ln, width, opt):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, SEA_LN=ln, SEA_DP_WIDTH=width, SEA_DP_OPT=opt)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", "scripts/dp_train_check.py"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "DP_TRAIN_OK" in r.stdout


@pytest.mark.parametrize("ln,width,opt", [("adaln", "small", "torch"), ("ln", "small", "fused"), ("adaln", "full", "fused")])
def test_dp_training_two_ranks_sharing_one_gpu(cuda, ln, width, opt):
    """The same check on ANY box, a 1-GPU one included: two ranks share cuda:0 and exchange over gloo (NCCL refuses two
    ranks on one device).  Everything on the device is the product path — backward with its per-group events, bucket
    planner, overlapped side-stream exchange, bf16 gradient twin, twin-reading fused AdamW; only the collective's
    transport differs from the NCCL cases above."""
    env = dict(os.environ, SEA_LN=ln, SEA_DP_WIDTH=width, SEA_DP_OPT=opt, SEA_DP_BACKEND="gloo")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29521", "scripts/dp_train_check.py"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "DP_TRAIN_OK" in r.stdout


def test_bench_two_gpus_weak_scaling_line(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _torchrun(2, ["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--rollout", "40"], 29519)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["value"] > 0
