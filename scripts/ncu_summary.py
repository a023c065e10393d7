"""one-screen summary of an ncu --set full report: python scripts/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
for d in data:
    print("kernel:", d[hdr.index("Kernel Name")][:90], "grid", d[hdr.index("Grid Size")], "block", d[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if h in KEYS or ("warp_issue_stalled" in h and h.endswith("per_warp_active.pct") and float(d[i] or 0) > 8) or "subpipe_dmma" in h:
            print(f"   {h} = {d[i]} {units[i]}")
