#!/bin/bash
# prints: kernel  registers  stack  spill-stores  spill-loads   from the *.ptxas.log files of the last build
cd "$(dirname "$0")"
python3 - <<'PY'
import glob, re
for f in sorted(glob.glob("*.ptxas.log")):
    t = open(f).read()
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads\n.*?Used (\d+) registers", t):
        print(f"{m.group(1)[:110]:110s} regs={m.group(5)} stack={m.group(2)} spill_st={m.group(3)} spill_ld={m.group(4)}")
PY
