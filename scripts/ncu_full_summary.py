"""Transposed summary (one column per launch) of selected metrics of an `ncu --set full` report.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/x_raw.csv
    python scripts/ncu_full_summary.py /tmp/x_raw.csv > profiles/x_ncu_full.csv
"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr, units, data = rows[0], rows[1], rows[2:]
iname = hdr.index("Kernel Name")
out = csv.writer(sys.stdout)
out.writerow(["metric"] + [r[iname].replace("sea::<", "").replace("void ", "") for r in data])
for m in WANT:
    if m in hdr:
        i = hdr.index(m)
        out.writerow([f"{m} [{units[i]}]"] + [r[i] for r in data])
