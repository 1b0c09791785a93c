#!/usr/bin/env python
"""Executed warp-instructions per CUDA source line, split into FP64-pipe and other opcodes, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`; counts are divided by `iters` (e.g. the number of
warp-level candidate iterations of the launch) so that they read as instructions per iteration.
    python tools/ncu_line_mix.py gpurun_out/prof_source_cuda.csv [iters] [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
iters = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
FP64 = {"DFMA", "DMUL", "DADD", "DSETP", "F2F", "MUFU"}
cur = curfile = None
per = collections.defaultdict(collections.Counter)
src, tot = {}, collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        curfile = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        continue
    if r[0] != "":
        try:
            cur = (curfile, int(r[0]))
        except ValueError:
            continue
        src[cur] = r[1][:70]
        continue
    if len(r) > 7 and r[3].strip() and r[3].strip() != "...":
        try:
            n = int(r[7])
        except ValueError:
            continue
        toks = r[3].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        per[cur][op] += n
        tot[op] += n
T = sum(tot.values())
F = sum(v for k, v in tot.items() if k in FP64)
print(f"total warp-instructions {T}  FP64-pipe {F} ({100 * F / T:.1f} %)  per iteration: total {T / iters:.1f} fp64 {F / iters:.1f} other {(T - F) / iters:.1f}")
print({k: round(v / iters, 1) for k, v in tot.most_common(30)})
lines = sorted(per.items(), key=lambda kv: -sum(v for k, v in kv[1].items() if k not in FP64))
print("--- lines by non-FP64 instructions per iteration ---")
for k, v in lines[:top]:
    o = sum(c for op, c in v.items() if op not in FP64)
    f = sum(c for op, c in v.items() if op in FP64)
    print(f"{k[0][:20]:20s}:{k[1]:4d} other {o / iters:7.1f} fp64 {f / iters:7.1f}  {dict((a, round(b / iters, 1)) for a, b in v.most_common(5))} | {src[k][:60]}")
