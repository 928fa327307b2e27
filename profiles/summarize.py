#!/usr/bin/env python
"""One line per `ncu --set full` raw-page CSV (ncu -i X.ncu-rep --page raw --csv): duration, DRAM bytes,
instruction count, pipe utilisation, issue activity, registers, shared-memory wavefronts.
usage: summarize.py file_raw.csv [...]"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__inst_executed.sum", "Minst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smemMwf"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid")]


def val(row, units, i):
    try:
        v = float(row[i])
    except ValueError:
        return None
    u = units[i]
    scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u)
    return v * scale if scale else v


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        print("%-28s (empty)" % path.split("/")[-1])
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        out = []
        for key, nm in KEYS:
            if key in hdr:
                v = val(r, units, hdr.index(key))
                if v is None:
                    continue
                if nm in ("Minst", "smemMwf"):
                    v /= 1e6
                out.append("%s %.4g" % (nm, v))
        name = r[hdr.index("Kernel Name")][:44] if "Kernel Name" in hdr else "?"
        print("%-24s %-44s %s" % (path.split("/")[-1].replace("_raw.csv", ""), name, "  ".join(out)))
