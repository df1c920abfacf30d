#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page raw --csv` dumps into a compact table (one row per captured launch).
usage: python scripts/ncu_summary.py gpurun_out/*_raw.csv > profiles/ncu_summary_rNN.md"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "tc_inst%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return v * mult


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


print("| file | kernel | time us | dram rd MB | dram wr MB | dram GB/s | dram% | tensor% | tc_inst% | fp64% | xu% | warps% | issue% | regs | grid | block |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].split("(")[0][-48:]
        def g(k):
            i = idx.get(k)
            return (r[i], units[i]) if i is not None and r[i] != "" else (None, None)
        t, tu = g("gpu__time_duration.sum")
        rd, ru = g("dram__bytes_read.sum")
        wr, wu = g("dram__bytes_write.sum")
        t_us = to_us(t, tu) if t else float("nan")
        rd_b = to_bytes(rd, ru) if rd else float("nan")
        wr_b = to_bytes(wr, wu) if wr else float("nan")
        out = [path.split("/")[-1].replace("_raw.csv", ""), name, f"{t_us:.1f}", f"{rd_b / 1e6:.1f}", f"{wr_b / 1e6:.1f}",
               f"{(rd_b + wr_b) / t_us / 1e3:.0f}"]
        for k, _ in COLS[3:]:
            v, _u = g(k)
            out.append(v if v is not None else "-")
        print("| " + " | ".join(out) + " |")
