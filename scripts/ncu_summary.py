"""Print the headline metrics of every kernel in an .ncu-rep (read offline with `ncu -i`). Usage: ncu_summary.py rep [substr...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp", "smsp__warp_issue_stalled", "lts__t_bytes.sum", "local_load", "local_store", "smsp__inst_executed_op_local",
        "l1tex__t_bytes_pipe_lsu_mem_local", "smsp__pcsamp_warps_issue_stalled", "sm__inst_executed_pipe", "smsp__average_warps_issue_stalled"]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:150])
    for i, h in enumerate(hdr):
        if any(k in h for k in KEYS + extra) and r[i] not in ("", "0", "0.000000"):
            print(f"  {h:90s} {units[i]:14s} {r[i]}")
