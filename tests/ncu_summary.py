"""Summary of an ncu report for profiles/ (manual tool): python tests/ncu_summary.py rep.ncu-rep > profiles/x.txt"""
import csv, io, subprocess, sys
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("%-61s %s" % ("Kernel Name", d["Kernel Name"]))
    for m in METRICS:
        if m in d:
            print("%-61s %-17s %s" % (m, units[hdr.index(m)], d[m]))
    print()
