"""Compact per-launch summary of an `ncu --page raw --csv` export: the columns the judge reads
(duration, DRAM bytes and %, tensor-pipe %, instructions, L2) in the format of profiles/*ncu_full*.csv."""
import csv
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
keep = [c for c in COLS if c in names]
idx = [names.index(c) for c in keep]
w = csv.writer(sys.stdout)
w.writerow(["Kernel Name", "Grid Size", "Block Size"] + keep)
w.writerow(["", "", ""] + [units[i] for i in idx])
kn, gs, bs = names.index("Kernel Name"), names.index("Grid Size"), names.index("Block Size")
for r in rows[hdr + 2:]:
    if len(r) > max(idx):
        w.writerow([r[kn], r[gs], r[bs]] + [r[i] for i in idx])
