#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one row per kernel launch with the counters the
roofline report uses (duration, DRAM bytes, L2 bytes, occupancy, issue rate, top stall)."""
import csv
import sys

KEYS = [
    ("dur_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("l2_MB", "lts__t_bytes.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("occ_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("ipc", "sm__inst_executed.avg.per_cycle_active"),
    ("inst", "smsp__inst_executed.sum"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("mem_pct", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
]


def to_mb(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit.lower(), 1.0)


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | " + " | ".join(k for k, _ in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    tot = 0.0
    for n, r in enumerate(rows[2:]):
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("eims::", "").replace("tc::", "")
        out = []
        for k, m in KEYS:
            if m not in col or r[col[m]] == "":
                out.append("-")
                continue
            v, u = r[col[m]], units[col[m]]
            if k.endswith("_MB"):
                out.append(f"{to_mb(v, u):.2f}")
            elif k == "dur_us":
                d = to_us(v, u)
                tot += d
                out.append(f"{d:.2f}")
            else:
                try:
                    out.append(f"{float(v.replace(',', '')):.4g}")
                except ValueError:
                    out.append(v)
        print(f"| {n} | {name} | " + " | ".join(out) + " |")
    print(f"\nsum of durations: {tot:.1f} us over {len(rows) - 2} launches")


if __name__ == "__main__":
    main(sys.argv[1])
