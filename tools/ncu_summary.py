#!/usr/bin/env python3
"""Summarise an `ncu -i X.ncu-rep --page raw --csv` export (one row per captured launch, one column per metric) into a
markdown table of the metrics the profiling recipe asks for.   usage: tools/ncu_summary.py raw.csv [title] > profiles/x.md"""
import csv
import sys

KEYS = [
    ("time us", "gpu__time_duration.sum", 1e-3, "{:.1f}"),
    ("DRAM read MB", "dram__bytes_read.sum", None, "{:.1f}"),
    ("DRAM write MB", "dram__bytes_write.sum", None, "{:.1f}"),
    ("DRAM % peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1, "{:.1f}"),
    ("tensor pipe % active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
    ("fma pipe % active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
    ("SM busy %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1, "{:.1f}"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1, "{:.1f}"),
    ("regs", "launch__registers_per_thread", 1, "{:.0f}"),
    ("smem KB/block", "launch__shared_mem_per_block_dynamic", None, "{:.0f}"),
]


def to_bytes(val, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(val) * scale


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def col(metric):
        for i, n in enumerate(hdr):
            if n == metric or n.endswith("." + metric):
                return i
        return None

    print(f"# {title}\n")
    print("ncu --set full --clock-control none, one row per captured launch (cold-cache, serialised: compare shares and ratios, not absolutes)\n")
    print("| kernel | grid x block | " + " | ".join(k for k, *_ in KEYS) + " | DRAM GB/s |")
    print("|---|---|" + "---|" * (len(KEYS) + 1))
    ik, ig, ib = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    for r in data:
        if len(r) != len(hdr):
            continue
        cells, t_us, dram = [], None, 0.0
        for label, metric, scale, fmt in KEYS:
            i = col(metric)
            if i is None or r[i] in ("", "n/a"):
                cells.append("-")
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if label == "time us":
                v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
                t_us = v
            elif "MB" in label:
                v = to_bytes(v, u) / 1e6
                dram += v
            elif "KB" in label:
                v = to_bytes(v, u) / 1e3
            cells.append(fmt.format(v))
        gbs = f"{dram / t_us * 1e3:.0f}" if t_us else "-"
        name = r[ik].split("(")[0][-70:]
        print(f"| `{name}` | {r[ig]} x {r[ib]} | " + " | ".join(cells) + f" | {gbs} |")


if __name__ == "__main__":
    main()
