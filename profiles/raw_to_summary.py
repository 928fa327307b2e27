#!/usr/bin/env python
"""`ncu --page raw --csv` (one column per metric) -> the 3-column summary (metric, unit, value per launch)
bench.py reads for `roofline.traffic`.   usage: raw_to_summary.py in_raw.csv out_summary.csv [launch]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
r = rows[2 + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)]
with open(sys.argv[2], "w", newline="") as fh:
    w = csv.writer(fh)
    for h, u, v in zip(hdr, units, r):
        if h in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"):
            continue
        w.writerow([h, u, v])
