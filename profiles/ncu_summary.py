#!/usr/bin/env python
"""Key metrics of every kernel in an ncu report, one block per launch:
    python profiles/ncu_summary.py gpurun_out/X.ncu-rep > profiles/rNN/X_summary.txt"""
import csv, subprocess, sys, io

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:110])
    vals = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            vals[w] = (r[i], units[i])
            print(f"  {w:66s} {r[i]:>18s} {units[i]}")
    try:
        f = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        t = {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}
        b = sum(float(vals[k][0].replace(",", "")) * f[vals[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        d = float(vals["gpu__time_duration.sum"][0].replace(",", "")) * t[vals["gpu__time_duration.sum"][1]]
        print(f"  {'-> DRAM traffic per launch':66s} {b / 1e9:18.3f} GB   = {b / d / 1e9:.0f} GB/s")
    except Exception as e:
        print("  (no traffic figure:", e, ")")
