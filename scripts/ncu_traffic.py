"""DRAM traffic of the GEMM kernel over ONE WHOLE rollout, from an ncu pass, next to the algorithmic bytes.

    # on the GPU box (one pass, three metrics; the plain run first, as the recipe requires):
    python scripts/rollout_one.py > gpurun_out/rollout_one.log 2>&1 && \
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:gemm_bf16 --csv --log-file gpurun_out/r2_gemm_dram.csv python scripts/rollout_one.py
    # here:
    python scripts/ncu_traffic.py gpurun_out/r2_gemm_dram.csv > profiles/r2_gemm_traffic.json

rollout_one.py runs the eager rollout twice; the second half of the launches is the measured rollout.
Algorithmic bytes of a launch = every operand read once + every output written once (bf16 operands, the
epilogue's residual read and fp32 / bf16 stores), summed over the problems grouped into the launch.
"""
import collections
import csv
import gzip
import json
import sys

E, H, Dd, V, B, R = 1024, 8192, 512, 2, 32, 100


def gemm_list(M, t):
    """(N, K, groups, output+residual bytes per element) of every GEMM launch of one inference forward at
    prefix length t (time-invariant condition: the AdaLN cond_mlp GEMMs run at t <= 2 only, on B rows)."""
    L = []
    if t <= 2:
        Mc = B if t == 2 else M
        L += [(Mc, 2 * E, 2 * E, 4, 4), (Mc, 2 * Dd, 2 * Dd, 2, 4), (Mc, 2 * E, 2 * E, 2, 4)]
    L += [(M, 3 * E, E, 2, 2), (M, E, E, 2, 10), (M, Dd, E, 2, 4), (M, Dd, Dd, 4, 2),
          (M, Dd, Dd, 1, 4), (M, E, Dd, 1, 10), (M, Dd, E, 1, 4),
          (M, 2 * Dd, Dd, 1, 2), (M, Dd, Dd, 1, 4), (M, E, Dd, 1, 10),
          (M, H, E, 2, 2), (M, E, H, 2, 6), (M, E, E, 2, 4)]
    return L


def algorithmic_bytes():
    tot, n = 0.0, 0
    for t in range(1, R + 1):
        for (M, N, K, g, ob) in gemm_list(B * t, t):
            tot += g * (2.0 * M * K + 2.0 * N * K + float(ob) * M * N)
            n += 1
    return tot, n


def main(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [l for l in f if l.startswith('"')]
    by_id = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = by_id.setdefault(r["ID"], {})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1,
              "usecond": 1e3, "msecond": 1e6}.get(u, 1)
        d[r["Metric Name"]] = v
    items = list(by_id.values())
    half = items[len(items) // 2:]          # second rollout
    dram = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in half)
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in half)
    t_ns = sum(d.get("gpu__time_duration.sum", 0) for d in half)
    alg, n_alg = algorithmic_bytes()
    out = {"dram_bytes_per_launch": dram / len(half), "algorithmic_bytes_per_launch": alg / n_alg,
           "dram_read_bytes_per_launch": rd / len(half), "launches": len(half), "launches_expected": n_alg,
           "dram_bytes_per_rollout": dram, "algorithmic_bytes_per_rollout": alg, "ratio": dram / alg,
           "gemm_time_ms_under_ncu": t_ns / 1e6,
           "note": "gemm_bf16_tn_kernel over one whole 100-step rollout (B=32, cylinder_flow) of the current kernels: ncu "
                   "dram__bytes_read.sum + dram__bytes_write.sum per launch (one pass, --clock-control none, cold-cache "
                   "serialised launches) vs operands-once + outputs-once; source " + path}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
