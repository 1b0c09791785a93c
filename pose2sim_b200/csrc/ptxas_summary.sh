#!/bin/bash
# prints: kernel  registers  stack  spill-stores  spill-loads   from the *.ptxas.log files
cd "$(dirname "$0")"
for f in *.ptxas.log; do
  awk '/Compiling entry function/ {name=$6} /bytes stack frame/ {stack=$1; ss=$5; sl=$9} /Used [0-9]+ registers/ {printf "%-90s regs=%s stack=%s spill_st=%s spill_ld=%s\n", name, $5, stack, ss, sl}' "$f"
done
