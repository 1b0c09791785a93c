#!/usr/bin/env python
"""Full SASS opcode histogram (static) of one kernel of a built object, from `cuobjdump -sass` — the evidence for which
instructions the binary contains (UBLKCP / SYNCS = TMA bulk copies + transaction barriers, CREDUX = redux.sync, DFMA ...).
    python tools/sass_histogram.py pose2sim_b200/csrc/p2s_triangulate.o 'triangulate_kernelILi8ELi0ELb0ELb1ELb0' > profiles/..."""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    ops = collections.Counter()
    for line in b.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            ops[m.group(1)] += 1
    base = collections.Counter()
    for k, v in ops.items():
        base[k.split(".")[0]] += v
    print(f"kernel {name}\ninstructions {sum(ops.values())}")
    print("--- by base opcode ---")
    for k, v in base.most_common():
        print(f"{k:16s} {v}")
    print("--- with modifiers ---")
    for k, v in ops.most_common():
        print(f"{k:40s} {v}")
    break
else:
    sys.exit(f"no kernel matching {pat}")
