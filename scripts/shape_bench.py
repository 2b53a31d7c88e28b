"""Per-shape microbenchmarks of the kernels on the temporal hot path (CUDA events, L2-cold inputs
rotated through a pool larger than L2).  Prints one line per shape.

    python scripts/shape_bench.py gemm            # every Linear shape of both configs at several M
    python scripts/shape_bench.py attn            # attention fwd / bwd at several (B, T, hd)
    python scripts/shape_bench.py train           # fwd+bwd train step of both configs + breakdown
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sea_b200 import ops  # noqa: E402
from sea_b200._lib import lib  # noqa: E402

dev = torch.device("cuda")


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def gemm_case(M, N, K, groups, bn=0, pool=6):
    As = [[torch.randn(M, K, device=dev).bfloat16() for _ in range(groups)] for _ in range(pool)]
    Ws = [[torch.randn(N, K, device=dev).bfloat16() * 0.02 for _ in range(groups)] for _ in range(pool)]
    Os = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(groups)]
    it = [0]

    def fn():
        k = it[0] % pool
        it[0] += 1
        probs = [ops.gemm_problem(As[k][g], Ws[k][g], out_bf16=Os[g]) for g in range(groups)]
        ops.gemm_bf16_tn(probs, M, N, K)

    lib.sea_gemm_force_tile_n(bn)
    us = timeit(fn)
    lib.sea_gemm_force_tile_n(0)
    return us


def run_gemm():
    for cfg, E, H, Dd in (("cyl", 1024, 8192, 512), ("mp", 2048, 16384, 1024)):
        shapes = [("qkv", 3 * E, E, 2), ("sproj", E, E, 2), ("down", Dd, E, 2), ("cq", Dd, Dd, 1),
                  ("ckv", 2 * Dd, Dd, 1), ("cproj", Dd, Dd, 1), ("up", E, Dd, 1), ("mlp0", H, E, 2),
                  ("mlp3", E, H, 2), ("proj", E, E, 2)]
        if cfg == "cyl":
            shapes.append(("cond2E", 2 * E, 2 * E, 4))
        for M in (32, 256, 800, 1600, 3200):
            tot_us, tot_fl = 0.0, 0.0
            for name, N, K, g in shapes:
                best = None
                res = []
                for bn in (0, 64, 128, 192, 256):
                    us = gemm_case(M, N, K, g, bn)
                    res.append((bn, us))
                auto = res[0][1]
                bn_b, us_b = min(res[1:], key=lambda r: r[1])
                fl = 2.0 * M * N * K * g
                print(f"{cfg} M={M:5d} {name:7s} N={N:5d} K={K:5d} g={g} auto {auto:7.1f} us {fl/auto/1e6:7.1f} TF/s | "
                      f"best bn={bn_b} {us_b:7.1f} us {fl/us_b/1e6:7.1f} TF/s | " +
                      " ".join(f"{bn}:{us:.1f}" for bn, us in res[1:]), flush=True)
                tot_us += auto
                tot_fl += fl
            print(f"== {cfg} M={M}: sum {tot_us:.1f} us, {tot_fl/tot_us/1e6:.1f} TF/s", flush=True)


def run_attn():
    for (B, T, nh, hd) in ((32, 100, 8, 128), (32, 100, 8, 64), (2, 399, 8, 128), (4, 199, 8, 256), (16, 399, 8, 128),
                           (8, 1024, 8, 128), (4, 2024, 8, 128), (4, 2024, 8, 64), (2, 2024, 8, 256)):
        M = B * T
        qkv = torch.randn(M, 3 * nh * hd, device=dev).bfloat16()
        q, k, v = qkv[:, : nh * hd], qkv[:, nh * hd: 2 * nh * hd], qkv[:, 2 * nh * hd:]
        us_f = timeit(lambda: ops.attention_fwd(q, k, v, nh, B=B, want_lse=True))
        o, lse = ops.attention_fwd(q, k, v, nh, B=B, want_lse=True)
        do = torch.randn_like(o)
        us_b = timeit(lambda: ops.attention_bwd(q, k, v, o, do, lse, nh, B=B), reps=5, warm=1)
        fl = 2.0 * B * nh * T * T * hd  # causal-useful fwd FLOPs (c = 1/2 of 4*T^2*hd)
        print(f"attn B={B} T={T} nh={nh} hd={hd}: fwd {us_f:8.1f} us {fl/us_f/1e6:7.1f} TF/s | "
              f"bwd {us_b:8.1f} us {2.5*fl/us_b/1e6:7.1f} TF/s", flush=True)


def run_attn_trace():
    """clock64 probes of CTA (0,0,0) of the two-tile forward kernel (the heaviest query block)."""
    import ctypes as C
    from sea_b200._lib import lib
    B, T, nh, hd = 4, 2024, 8, 128
    qkv = torch.randn(B * T, 3 * nh * hd, device=dev).bfloat16()
    q, k, v = qkv[:, : nh * hd], qkv[:, nh * hd: 2 * nh * hd], qkv[:, 2 * nh * hd:]
    buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.attention_fwd(q, k, v, nh, B=B)
    lib.sea_attention_debug_trace(C.c_void_p(buf.data_ptr()))
    ops.attention_fwd(q, k, v, nh, B=B)
    torch.cuda.synchronize()
    lib.sea_attention_debug_trace(C.c_void_p(0))
    t = buf.cpu().view(3, 64, 8)
    t0 = int(t[t > 0].min())
    n = (T + 127) // 128
    print("softmax group g, tile j: [s_full seen, ld done, max/rescale done, exp done, P stored+arrive] (cycles from first probe)")
    for j in range(n):
        for g in range(2):
            r = [int(v) - t0 if int(v) else -1 for v in t[g, j, :5]]
            m = [int(v) - t0 if int(v) else -1 for v in t[2, j, g * 3: g * 3 + 3]]
            print(f"j={j:2d} g={g}: sm {r}   mma [p_full seen, kv seen, issued] {m}")


def run_train():
    from sea_b200.rollout import profile
    from sea_b200.temporal import TemporalModel
    for cfg, E, ln, B, T in (("cylinder_flow", 1024, "adaln", 2, 399), ("multiphase_flow", 2048, "ln", 4, 199),
                             ("cylinder_flow", 1024, "adaln", 16, 399), ("multiphase_flow", 2048, "ln", 16, 199)):
        torch.manual_seed(42)
        m = TemporalModel(1, E, 8, 2024, 8, 0, 2, 2, 0.0, "sea", "learnable", "mlp", "add", 1, 1, True, ln).to(dev)
        m.train()
        opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.0)
        x = torch.randn(B, T, 2, E, device=dev)
        ib = torch.rand(B, 1, 1, device=dev).expand(B, T, 1).contiguous()
        tgt = torch.randn_like(x)

        def step(with_opt=True):
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.mse_loss(m(x, ib), tgt)
            loss.backward()
            if with_opt:
                opt.step()

        us_fb = timeit(lambda: step(False), reps=5, warm=2)
        us_all = timeit(lambda: step(True), reps=5, warm=2)
        from sea_b200.optim import AdamW as FusedAdamW
        opt = FusedAdamW(m.parameters(), lr=1e-4, weight_decay=0.0, engine=m.engine())
        us_fused = timeit(lambda: step(True), reps=5, warm=2)
        with torch.no_grad():
            m.eval()
            us_f = timeit(lambda: m(x, ib), reps=5, warm=2)
            m.train()
        with profile() as prof:
            step(False)
            torch.cuda.synchronize()
        ps = prof.summary
        M = B * T
        from bench import fwd_flops  # noqa
        Hh, Dd = 8 * E, E // 2
        fl = fwd_flops(B, T, E=E, H=Hh, Dd=Dd, adaln=(ln == "adaln"))
        print(f"train {cfg} B={B} T={T} M={M}: fwd {us_f:.0f} us ({fl/us_f/1e6:.0f} TF/s) fwd+bwd {us_fb:.0f} us "
              f"({3*fl/us_fb/1e6:.0f} TF/s) +torch AdamW {us_all:.0f} us / +fused AdamW {us_fused:.0f} us | breakdown ms " +
              " ".join(f"{k}:{v['ms']:.2f}/{v['launches']}" for k, v in ps.items()) +
              f" | gemm {ps['gemm']['work']/max(ps['gemm']['ms'],1e-9)/1e9:.0f} TF/s "
              f"attn {ps['attention']['work']/max(ps['attention']['ms'],1e-9)/1e9:.1f} TF/s "
              f"elem {ps['elementwise']['work']/max(ps['elementwise']['ms'],1e-9)/1e6:.0f} GB/s", flush=True)
        del m, opt


def run_rollout():
    """Per-step device time of the headline rollout (cylinder, B=32) as a function of the prefix length."""
    from bench import build_model, make_inputs
    B, R = 32, 100
    m = build_model("bf16").to(dev).eval()
    x0, ib = make_inputs(B, R, 1024, 2, 1234)
    x0, ib = x0.to(dev), ib.to(dev)
    eng = m.engine()
    from sea_b200.rollout import rollout
    for _ in range(2):
        rollout(m, x0, ib, R)
    torch.cuda.synchronize()
    eng.ib_time_invariant, eng.cond_reuse, eng._cond_valid = True, True, False
    seq = x0
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(R + 1)]
    import time
    host = []
    with torch.no_grad():
        evs[0].record()
        for i in range(R):
            t0 = time.perf_counter()
            out = m(seq, ib[:, : i + 1])
            seq = torch.cat((seq, out[:, -1:]), dim=1)
            host.append((time.perf_counter() - t0) * 1e6)
            evs[i + 1].record()
            if i % 10 == 9:
                torch.cuda.synchronize()   # keep the host from running ahead: host[] = pure issue time
    torch.cuda.synchronize()
    eng.ib_time_invariant, eng.cond_reuse, eng._cond_valid = False, False, False
    devt = [evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(R)]
    for i in range(0, R, 5):
        print(f"rollout t={i+1:3d} M={B*(i+1):5d}: device {devt[i]:7.1f} us  host-issue {host[i]:7.1f} us  launches {eng.last_launches}")
    print(f"sum device {sum(devt)/1e3:.2f} ms, sum host {sum(host)/1e3:.2f} ms")
    # graph replays
    from sea_b200.rollout import RolloutPlan
    plan = RolloutPlan(m, B, R, dev)
    plan.run(x0, ib)
    torch.cuda.synchronize()
    evs[0].record()
    for i, g in enumerate(plan.graphs):
        g.replay()
        evs[i + 1].record()
    torch.cuda.synchronize()
    devg = [evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(R)]
    for i in range(0, R, 5):
        print(f"graph   t={i+1:3d} M={B*(i+1):5d}: device {devg[i]:7.1f} us  nodes {plan.launches[i]}")
    print(f"sum graph device {sum(devg)/1e3:.2f} ms")
    # KV-cached incremental engine
    from sea_b200.rollout import CachedRolloutPlan
    cp = CachedRolloutPlan(m, B, R, dev, True)
    cp.run(x0, ib)
    torch.cuda.synchronize()
    evs[0].record()
    for i, g in enumerate(cp.graphs):
        g.replay()
        evs[i + 1].record()
    torch.cuda.synchronize()
    devc = [evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(R)]
    for i in range(0, R, 10):
        print(f"cached  t={i+1:3d}: device {devc[i]:7.1f} us  nodes {cp.launches[i]}")
    print(f"sum cached device {sum(devc)/1e3:.2f} ms")




def run_floor():
    """Per-node device time of CUDA-graph chains of identical small kernels (latency floor)."""
    def chain(fn, n=48, reps=20):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=side, capture_error_mode="relaxed"):
            for _ in range(n):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps / n * 1e3

    x = torch.randn(160, 1024, device=dev)
    w = torch.ones(1024, device=dev)
    y = torch.empty(160, 1024, device=dev)
    print(f"torch add_ [160x1024]           : {chain(lambda: y.add_(1.0)):6.2f} us/node")
    print(f"norm_fwd LN [160x1024]          : {chain(lambda: ops.norm_fwd(x, w, kind=0)):6.2f} us/node")
    g8 = torch.randn(32, 8, device=dev)
    for (M, N, K) in ((160, 512, 512), (160, 1024, 1024), (160, 3072, 1024), (160, 8192, 1024), (160, 1024, 8192),
                      (1600, 512, 512), (1600, 1024, 1024)):
        a = torch.randn(M, K, device=dev).bfloat16()
        b = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for st in (False, True):
            pr = [ops.gemm_problem(a, b, out_bf16=o, b_is_static=st)]
            print(f"gemm M={M} N={N} K={K} static={int(st)}: {chain(lambda: ops.gemm_bf16_tn(pr, M, N, K)):6.2f} us/node")
    for (B, T, nh, hd) in ((32, 5, 8, 128), (32, 5, 8, 64), (32, 100, 8, 128)):
        qkv = torch.randn(B * T, 3 * nh * hd, device=dev).bfloat16()
        q, k, v = qkv[:, : nh * hd], qkv[:, nh * hd: 2 * nh * hd], qkv[:, 2 * nh * hd:]
        print(f"attn B={B} T={T} hd={hd}: {chain(lambda: ops.attention_fwd(q, k, v, nh, B=B)):6.2f} us/node")
    for pdl in (0, 1):
        lib.sea_set_pdl(pdl)
        a = torch.randn(160, 512, device=dev).bfloat16()
        b = (torch.randn(512, 512, device=dev) * 0.02).bfloat16()
        o = torch.empty(160, 512, device=dev, dtype=torch.bfloat16)
        pr = [ops.gemm_problem(a, b, out_bf16=o, b_is_static=True)]
        print(f"pdl={pdl} gemm 160x512x512 chain: {chain(lambda: ops.gemm_bf16_tn(pr, 160, 512, 512)):6.2f} us/node")
    lib.sea_set_pdl(1)



def run_kblock():
    """Per-k-block time of one CTA column: graph chain of a K-heavy GEMM at forced tile widths."""
    import itertools
    def chain(fn, n=16, reps=10):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=side, capture_error_mode="relaxed"):
            for _ in range(n):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps / n * 1e3
    for (M, N, K) in ((128, 256, 8192), (128, 8192, 8192)):
        a = torch.randn(M, K, device=dev).bfloat16()
        b = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        pr = [ops.gemm_problem(a, b, out_bf16=o, b_is_static=True)]
        for bn, dbg in itertools.product((64, 128, 256), (0, 1, 2)):
            if bn > N:
                continue
            lib.sea_gemm_force_tile_n(bn)
            lib.sea_gemm_debug_probe(dbg)
            us = chain(lambda: ops.gemm_bf16_tn(pr, M, N, K))
            lib.sea_gemm_force_tile_n(0)
            lib.sea_gemm_debug_probe(0)
            ctas = ((M + 127) // 128) * ((N + bn - 1) // bn)
            kb = K // 64
            print(f"M={M} N={N} K={K} bn={bn} probe={dbg}: {us:7.2f} us, {ctas} CTAs, {us/kb*1e3:6.1f} ns/k-block, "
                  f"{(128+bn)*128/ (us/kb*1e3):6.1f} GB/s per CTA", flush=True)


def run_sk():
    """Stream-K vs data-parallel schedule of the same GEMM (graph chains of 8 launches, warm)."""
    import ctypes as C
    ws = torch.zeros(65536 + 148 * 2 * 128 * 256 * 4, dtype=torch.uint8, device=dev)

    def chain(fn, n=8, reps=10):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=side, capture_error_mode="relaxed"):
            for _ in range(n):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps / n * 1e3
    lib.sea_gemm_set_workspace(C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()))
    for (M, N, K, G) in ((32, 1024, 8192, 2), (800, 1024, 8192, 2), (1600, 1024, 8192, 2), (3200, 1024, 8192, 2),
                         (3200, 8192, 1024, 2), (800, 3072, 1024, 2), (3200, 1024, 1024, 2), (796, 2048, 16384, 2)):
        a = [torch.randn(M, K, device=dev).bfloat16() for _ in range(G)]
        b = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(G)]
        o = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(G)]
        pr = [ops.gemm_problem(a[i], b[i], out_bf16=o[i], b_is_static=True) for i in range(G)]
        fl = 2.0 * M * N * K * G
        row = []
        for bn in (0, 64, 128, 192, 256):
            for mode in (0, 2):
                lib.sea_gemm_force_tile_n(bn); lib.sea_gemm_stream_k(mode)
                us = chain(lambda: ops.gemm_bf16_tn(pr, M, N, K))
                row.append(f"bn{bn}/{'sk' if mode else 'dp'} {us:6.1f}us {fl/us/1e6:5.0f}TF")
        lib.sea_gemm_force_tile_n(0); lib.sea_gemm_stream_k(1)
        us = chain(lambda: ops.gemm_bf16_tn(pr, M, N, K))
        print(f"M={M} N={N} K={K} g={G}: auto {us:6.1f}us {fl/us/1e6:5.0f}TF | " + " | ".join(row), flush=True)
    lib.sea_gemm_set_workspace(None, C.c_size_t(0))


def run_patchify():
    """Mesh patchify / unpatch throughput (HBM-bound gather / scatter), 60k-cell mesh, 3 fields."""
    from sea_b200.patchify import DataPartitioner2D
    g = torch.Generator(device="cuda").manual_seed(1)
    N, F = 60_000, 3
    x = torch.rand(N, device=dev, generator=g) * 2.2
    y = torch.rand(N, device=dev, generator=g) * 0.41
    for S in (256, 2048):
        vars_ = [torch.randn(S, N, device=dev, generator=g) for _ in range(F)]
        part = DataPartitioner2D(x, y, device=dev)
        t_idx = timeit(lambda: part._build_index(), reps=5, warm=1)
        stacked = torch.stack(vars_, 0)
        fields = part.gather(vars_)
        us_g = timeit(lambda: part.gather(vars_), reps=10, warm=2) - timeit(lambda: torch.stack(vars_, 0), reps=10, warm=2)
        us_gp = timeit(lambda: part.gather(vars_, layout_pfc=True), reps=10, warm=2)
        us_s = timeit(lambda: part.scatter(fields), reps=10, warm=2)
        P, Cc = part.index_map_tensor.shape
        by_g = 4.0 * S * N * F + 4.0 * S * P * Cc * F
        print(f"patchify S={S} N={N} F={F} P={P} C={Cc}: index {t_idx:.0f} us | gather {us_g:.0f} us "
              f"({by_g/us_g/1e3:.0f} GB/s) | gather+stack [S,P,F,C] {us_gp:.0f} us | scatter {us_s:.0f} us ({by_g/us_s/1e3:.0f} GB/s)", flush=True)
        del vars_, fields, stacked


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
    if what == "attn_trace":
        run_attn_trace()
        sys.exit(0)
    {"gemm": run_gemm, "attn": run_attn, "train": run_train, "rollout": run_rollout, "floor": run_floor, "kblock": run_kblock, "sk": run_sk, "patchify": run_patchify}[what]()
