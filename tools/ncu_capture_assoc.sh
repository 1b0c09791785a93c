#!/bin/bash
# One `ncu --set full` capture (with source correlation) of the association kernel on the 8 cameras x 3 persons x 20 k
# frames shape.  Usage on the GPU box: bash tools/ncu_capture_assoc.sh <tag>
set -e
tag=${1:-assoc}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:associate_kernel --launch-skip ${2:-4} --launch-count 1 \
    -o gpurun_out/${tag} -f python tools/assoc_ab.py one > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_source_cuda.csv
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv
ls -la gpurun_out/${tag}*
