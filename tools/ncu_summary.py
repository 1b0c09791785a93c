#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> one (metric, unit, value per launch) row per metric — the form kept
under profiles/.  Usage:  ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/x.csv"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, launches = rows[hi], rows[hi + 1], rows[hi + 2:]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
for c, name in enumerate(names):
    if name in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Block Size", "Grid Size", "Device", "CC"):
        if name not in ("Block Size", "Grid Size"):
            continue
    w.writerow([name, units[c] if c < len(units) else ""] + [r[c] if c < len(r) else "" for r in launches])
