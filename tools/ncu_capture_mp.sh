#!/bin/bash
# One `ncu --set full` capture (with source correlation) of the multi-person matching kernel on the cfg4 shape
# (8 cameras x 6 persons, 48 detections per frame, 1000 frames), after the same command ran once without ncu.
# Usage on the GPU box: bash tools/ncu_capture_mp.sh <tag>     (P2S_MP_ONE_RESIDENT=1 in the environment: the round-2 launch)
set -e
tag=${1:-mp}
mkdir -p gpurun_out
python tests/perf/mp_bench.py 1000 > gpurun_out/${tag}_plain.json
ncu --set full --clock-control none --import-source on -k regex:mp_associate_kernel --launch-skip 1 --launch-count 1 \
    -o gpurun_out/${tag} -f python tests/perf/mp_bench.py 1000 > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv | python tools/ncu_summary.py > gpurun_out/${tag}_ncu_full.csv
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_source.csv
python tools/ncu_lines.py gpurun_out/${tag}_source.csv > gpurun_out/${tag}_source_lines.txt 2>/dev/null || true
ls -la gpurun_out/${tag}*
