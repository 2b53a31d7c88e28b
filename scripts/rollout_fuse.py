"""A/B: steps per CUDA graph in the prefix-recompute rollout plan (SEA_ROLLOUT_STEPS_PER_GRAPH; 0 = whole rollout).
    python scripts/rollout_fuse.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
import bench
from sea_b200 import rollout as R
dev = torch.device("cuda", 0)
m = bench.build_model("bf16").to(dev).eval()
x0, ib = bench.make_inputs(32, 100, 1024, 2, 0)
x0, ib = x0.to(dev), ib.to(dev)
out = R.rollout(m, x0, ib, 100, _view_ok=True)
ref = out.clone()
torch.cuda.synchronize()
ts = []
for _ in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = R.rollout(m, x0, ib, 100, _view_ok=True); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print("steps/graph", os.environ.get("SEA_ROLLOUT_STEPS_PER_GRAPH"), "median ms", round(ts[len(ts)//2], 3), "min", round(ts[0], 3),
      "checksum", float(out.double().abs().sum()))
''' % ROOT
for per in ("1", "5", "20", "0"):
    env = dict(os.environ, SEA_ROLLOUT_STEPS_PER_GRAPH=per)
    subprocess.run([sys.executable, "-c", CHILD], env=env, check=False)
