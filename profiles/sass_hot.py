#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output:
top instructions by stall samples, stall-reason totals, instruction mix.
usage: sass_hot.py file.csv [kernel_index]"""
import csv
import sys


def main(path, which=0, topn=40):
    rows = list(csv.reader(open(path)))
    heads = [i for i, r in enumerate(rows) if "Source" in r and len(r) > 5]
    hi = heads[min(which, len(heads) - 1)]
    end = heads[heads.index(hi) + 1] if heads.index(hi) + 1 < len(heads) else len(rows)
    H = rows[hi]
    data = [r for r in rows[hi + 1:end] if len(r) == len(H)]
    ai, si = H.index("Address"), H.index("Source")
    sa, ie = H.index("Warp Stall Sampling (All Samples)"), H.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_")]

    def num(x):
        try:
            return int(float(x))
        except ValueError:
            return 0
    tot = sum(num(r[sa]) for r in data)
    print("instructions: %d   stall samples: %d   warp-instructions executed: %d"
          % (len(data), tot, sum(num(r[ie]) for r in data)))
    for r in sorted(data, key=lambda r: -num(r[sa]))[:topn]:
        st = sorted(((H[i], num(r[i])) for i in stall_cols if num(r[i]) > 0), key=lambda kv: -kv[1])[:3]
        print("%6s %5.1f%% inst=%9d  %-64s %s" % (r[ai][-5:], 100.0 * num(r[sa]) / max(tot, 1), num(r[ie]),
                                                  r[si][:64], st))
    agg = {}
    for r in data:
        for i in stall_cols:
            agg[H[i]] = agg.get(H[i], 0) + num(r[i])
    print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:10])
    mix = {}
    for r in data:
        op = r[si].split()[0] if r[si].split() else "?"
        if op.startswith("@"):
            op = r[si].split()[1]
        op = op.split(".")[0]
        mix[op] = mix.get(op, 0) + num(r[ie])
    print("executed mix:", sorted(mix.items(), key=lambda kv: -kv[1])[:16])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
