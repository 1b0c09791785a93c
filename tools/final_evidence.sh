#!/bin/bash
# The per-tree evidence set kept under profiles/ (one B200):  bash tools/final_evidence.sh <tag> [full]
#   <tag>_pytest_gpu.log        pytest -m gpu
#   <tag>_bench_n1.json         default bench.py line            <tag>_bench_reference.json  the --impl reference arm
#   <tag>_launches.csv          ncu launch list of `bench.py --steps 2 --warmup 1 --no-cpu-baseline`
#                               (--metrics gpu__time_duration.sum --clock-control none), only after the plain run exited 0
#   full: <tag>_tri16_* / <tag>_deep_* — one `ncu --set full` capture each of the 16-camera search kernel and of
#         deep_search_kernel on the cfg3 shard (tools/ncu_capture.sh)
# Everything lands in gpurun_out/; nothing printed under ncu is a bench value.
set -o pipefail
tag=${1:-final}; full=${2:-}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/${tag}_pytest_gpu.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
if python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_bench_short.json 2>&1; then
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu list rc=$?"
fi
if [ -n "$full" ]; then
    bash tools/ncu_capture.sh ${tag}_tri16 cfg3 > gpurun_out/${tag}_tri16_capture.log 2>&1; echo "ncu full rc=$?"
    python tools/ncu_summary.py < gpurun_out/${tag}_tri16_raw.csv > gpurun_out/${tag}_triangulate16_ncu_full.csv
    rm -f gpurun_out/${tag}_tri16.ncu-rep gpurun_out/${tag}_tri16_source.csv
    bash tools/ncu_capture.sh ${tag}_deep cfg3 deep_search_kernel > gpurun_out/${tag}_deep_capture.log 2>&1; echo "ncu deep rc=$?"
    python tools/ncu_summary.py < gpurun_out/${tag}_deep_raw.csv > gpurun_out/${tag}_deep_search_ncu_full.csv
    rm -f gpurun_out/${tag}_deep.ncu-rep gpurun_out/${tag}_deep_source.csv
fi
python tools/show_bench.py gpurun_out/${tag}_bench_n1.json 2>/dev/null | head -30
