#!/usr/bin/env python3
"""profiles/ncu_<tag>_all_kernels.csv from an `ncu --set full` report of tools/gpu_prof_all.py: one row per launch with the
columns the earlier rounds used (time, launch shape, registers, shared memory, occupancy limits, pipe and issue activity,
instructions, DRAM bytes, L2 throughput).   python tools/ncu_all_summary.py report.ncu-rep out.csv "comment"."""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
comment = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = ["ID", "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [hdr.index(c) if c in hdr else -1 for c in cols]
with open(out, "w", newline="") as f:
    if comment:
        f.write("# " + comment + "\n")
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[i] if i >= 0 else "" for i in idx])
    for r in data:
        w.writerow([r[i][:90] if i >= 0 else "" for i in idx])
print(len(data), "launches ->", out)
