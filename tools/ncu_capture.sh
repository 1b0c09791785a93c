#!/bin/bash
# One `ncu --set full` capture (with source correlation) of the lean triangulation kernel on cfg2, after the same
# command has run once without ncu.  Usage on the GPU box: bash tools/ncu_capture.sh <tag> [workload] [kernel regex]
# Leaves gpurun_out/<tag>.ncu-rep, <tag>_raw.csv, <tag>_source.csv, <tag>_source_cuda.csv.
set -e
tag=${1:-prof}; wl=${2:-cfg2}; kern=${3:-triangulate_kernel}
mkdir -p gpurun_out
python tools/kernel_ab.py one $wl 3 > gpurun_out/${tag}_plain.json
ncu --set full --clock-control none --import-source on -k regex:$kern --launch-skip 3 --launch-count 1 \
    -o gpurun_out/${tag} -f python tools/kernel_ab.py one $wl 3 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_source_cuda.csv
ls -la gpurun_out/${tag}*
