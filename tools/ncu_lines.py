#!/usr/bin/env python
"""Per CUDA-source-line shares of warp-stall samples and executed instructions from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (needs -lineinfo).  Usage:
    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv; python tools/ncu_lines.py /tmp/src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
sec, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0])
txt = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        sec, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Line No":
        hdr = {}
        for i, h in enumerate(r):
            hdr.setdefault(h, i)
        continue
    if hdr is None or sec is None:
        continue
    try:
        ln, s, ie = int(r[0]), int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0)
        xw = int(r[hdr["L1 Wavefronts Shared Excessive"]] or 0)
    except Exception:
        continue
    a = agg[(sec, ln)]
    a[0] += s; a[1] += ie; a[2] += xw
    txt[(sec, ln)] = r[1].strip()[:100]
ts = sum(v[0] for v in agg.values())
ti = sum(v[1] for v in agg.values())
print(f"samples {ts}  executed warp-instructions {ti}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:20s}:{k[1]:4d} samples {100 * v[0] / ts:5.2f}%  inst {100 * v[1] / ti:5.2f}%  excess smem wavefronts {v[2]:8d} | {txt[k]}")
