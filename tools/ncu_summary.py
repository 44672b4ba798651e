#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per profiled launch (duration, DRAM traffic, pipe /
memory throughput, occupancy).  Usage: ncu -i X.ncu-rep --page raw --csv > X_raw.csv; tools/ncu_summary.py X_raw.csv"""
import csv
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    rows = list(csv.reader(open(path)))
    head, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(head)}
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}

    def col(row, k):
        return num(row[idx[k]]) if k in idx else float("nan")

    dram_pct = next((k for k in ("dram__throughput.avg.pct_of_peak_sustained_elapsed",
                                 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") if k in idx), None)
    print("%-36s %6s %4s %4s %8s %8s %8s %6s %6s %6s %6s %6s %6s" % (
        "kernel", "grid", "blk", "reg", "dur_us", "rd_MB", "wr_MB", "sm%", "dram%", "l1%", "l2%", "occ%", "l2hit"))
    for row in rows[2:]:
        t = col(row, "gpu__time_duration.sum") * tscale.get(units[idx["gpu__time_duration.sum"]], 1.0)
        rd = col(row, "dram__bytes_read.sum") * scale.get(units[idx["dram__bytes_read.sum"]], 1.0)
        wr = col(row, "dram__bytes_write.sum") * scale.get(units[idx["dram__bytes_write.sum"]], 1.0)
        name = row[idx["Kernel Name"]].replace("<unnamed>::", "")
        print("%-36s %6s %4s %4s %8.1f %8.2f %8.2f %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f" % (
            name[:36], row[idx["launch__grid_size"]], row[idx["launch__block_size"]],
            row[idx["launch__registers_per_thread"]], t, rd, wr,
            col(row, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), col(row, dram_pct) if dram_pct else float("nan"),
            col(row, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            col(row, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            col(row, "sm__warps_active.avg.pct_of_peak_sustained_active"), col(row, "lts__t_sector_hit_rate.pct")))


if __name__ == "__main__":
    main(sys.argv[1])
