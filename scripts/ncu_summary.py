"""Extract a per-kernel summary (one row per captured launch) from an .ncu-rep into a small CSV for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r1_ncu_prof_x_summary.csv
"""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__cluster_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second"]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
cols = [k for k in KEEP if k in h] + [k for k in h if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[h.index(k)] for k in cols])
    for r in rows[2:]:
        w.writerow([r[h.index(k)] for k in cols])
print("wrote", sys.argv[2], len(rows) - 2, "launches")
