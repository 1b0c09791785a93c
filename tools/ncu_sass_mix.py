#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): executed warp-instructions per opcode,
average active lanes, stall-sample shares, and the hottest instructions.  Usage:
    ncu -i gpurun_out/prof.ncu-rep --page source --csv > /tmp/src.csv; python tools/ncu_sass_mix.py /tmp/src.csv [top]"""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:                      # first launch only (the header repeats per launch)
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) == len(hdr):
        data.append(r)
ops = collections.Counter()
lanes = collections.Counter()
samples = collections.Counter()
tot_i = tot_s = 0
stall_cols = [h for h in hdr if h.startswith("stall_")]
stalls = collections.Counter()
recs = []
for n, r in enumerate(data):
    src = r[col["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    base = op.split(".")[0]
    if base in ("MUFU", "F2F", "LDS", "STS", "LDG", "STG"):
        base = ".".join(op.split(".")[:2]) if base == "MUFU" else base
    ie = int(r[col["Instructions Executed"]] or 0)
    te = int(r[col["Thread Instructions Executed"]] or 0)
    sm = int(r[col["# Samples"]] or 0)
    ops[base] += ie
    lanes[base] += te
    samples[base] += sm
    tot_i += ie
    tot_s += sm
    for h in stall_cols:
        stalls[h] += int(r[col[h]] or 0)
    recs.append((sm, ie, n, src, {h: int(r[col[h]] or 0) for h in stall_cols if int(r[col[h]] or 0)}))
print(f"SASS instructions {len(data)}  executed warp-inst {tot_i}  samples {tot_s}")
for op, n in ops.most_common(top):
    print(f"{op:22s} {n:12d} {100.0 * n / tot_i:6.2f}%  samples {samples[op]:7d} ({100.0 * samples[op] / max(tot_s, 1):5.1f}%)  lanes {lanes[op] / max(n, 1):5.1f}")
ts = sum(stalls.values())
print({h: round(100.0 * v / max(ts, 1), 1) for h, v in stalls.most_common(12)})
print("hottest instructions:")
for sm, ie, n, src, st in sorted(recs, reverse=True)[:top]:
    print(f"{n:5d} {src[:70]:70s} samples {sm:6d} exec {ie:9d} {sorted(st.items(), key=lambda kv: -kv[1])[:3]}")
