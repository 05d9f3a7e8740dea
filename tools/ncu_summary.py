#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/: one row per captured launch."""
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["%s [%s]" % (k, units[i]) if units[i] else k for k, i in cols])
        for r in rows[2:]:
            w.writerow([r[i] for _, i in cols])
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
