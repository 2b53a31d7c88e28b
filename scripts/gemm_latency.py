"""Where does a SMALL-M GEMM launch spend its time?  (the rollout's short prefixes are latency-bound;
every GEMM node costs 9-15 us however little it does.)

For every Linear shape of the cylinder_flow forward at M = 32 / 320 / 960 rows:
  * chain : 40 back-to-back launches on one stream (PDL on), CUDA events -> us per launch in a dependent chain;
  * trace : %globaltimer of CTA 0 at 8 hand-off points of ONE launch (sea_gemm_debug_trace), as deltas in ns:
            entry->prologue | ->pdl wait passed | ->first stage landed | ->last MMA committed | ->epilogue sees acc |
            ->epilogue stored | ->exit
  * the tile / grid / stream-K decision of the launch.

    python scripts/gemm_latency.py > gpurun_out/gemm_latency.txt

The trace columns need a library built with the probe points compiled in (they are off in the product build because a
`lane == 0` test inside the MMA warp's loop costs the uniform-datapath issue):
    touch sea_b200/csrc/gemm.cu && NVCC_APPEND_FLAGS=-DSEA_GEMM_TRACE python -c "import __graft_entry__ as g; g.build()"
Result of this round: profiles/r2_gemm_latency.md.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200 import ops  # noqa: E402
from sea_b200._lib import check, lib  # noqa: E402

dev = torch.device("cuda")
check(lib.sea_init(0), "init")
E, H, Dd = 1024, 8192, 512
SHAPES = [("qkv", 3 * E, E, 2), ("sproj+res", E, E, 2), ("down", Dd, E, 2), ("cross q|k|v x4", Dd, Dd, 4),
          ("ckv", 2 * Dd, Dd, 1), ("cproj", Dd, Dd, 1), ("up", E, Dd, 1), ("mlp0", H, E, 2), ("mlp3", E, H, 2),
          ("proj", E, E, 2)]
ws = torch.zeros(65536 + 148 * 2 * 128 * 256 * 4, dtype=torch.uint8, device=dev)
check(lib.sea_gemm_set_workspace(C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel())), "ws")
trace = torch.zeros(8, dtype=torch.int64, device=dev)
cfg = (C.c_int * 4)()

print(f"{'shape':18s} {'M':>5s} {'N':>6s} {'K':>6s} g | bn grid upc tiles | chain us | trace ns: prolog pdlwait 1st-stage mainloop "
      f"acc-seen epilogue exit | total")
lib.sea_gemm_debug_probe(int(os.environ.get('SEA_PROBE', '0')))
for M in (32, 320, 960, 3200):
    for name, N, K, g in SHAPES:
        A = [torch.randn(M, K, device=dev).bfloat16() for _ in range(g)]
        W = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(g)]
        O = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(g)]

        def fn():
            probs = [ops.gemm_problem(A[i], W[i], out_bf16=O[i], b_is_static=True) for i in range(g)]
            ops.gemm_bf16_tn(probs, M, N, K)

        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            fn()
        e1.record()
        torch.cuda.synchronize()
        chain = e0.elapsed_time(e1) / 40 * 1e3
        lib.sea_gemm_last_config(cfg)
        lib.sea_gemm_debug_trace(C.c_void_p(trace.data_ptr()))
        fn()
        torch.cuda.synchronize()
        lib.sea_gemm_debug_trace(None)
        t = trace.cpu().tolist()
        d = [t[i + 1] - t[i] for i in range(7)]
        print(f"{name:18s} {M:5d} {N:6d} {K:6d} {g} | {cfg[0]:3d} {cfg[1]:4d} {cfg[2]:3d} {cfg[3]:5d} | {chain:8.2f} | "
              + " ".join(f"{x:7d}" for x in d) + f" | {t[7] - t[0]:6d}")
    print()
