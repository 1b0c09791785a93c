#!/bin/bash
# Same-box A/B of the deep-level kernel's shape (P2S_DEEP_CLUSTER / P2S_DEEP_GRID_MULT builds under pose2sim_b200/ab/,
# tools/kernel_ab.py build) and of the parking threshold, on the cfg3 shard, after the full GPU suite on the default build;
# then the default bench line and a short fuzz.     bash tools/deep_grid_ab.sh <tag>
tag=${1:-deep}
mkdir -p gpurun_out
timeout 170 python -m pytest tests -q -m gpu -x > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/${tag}_pytest_gpu.log
timeout 60 python tools/kernel_ab.py run cfg3 20 > gpurun_out/${tag}_kernel_ab_deep.jsonl 2>&1
for dm in 1000 400; do
    P2S_DEEP_MIN=$dm timeout 30 python tools/kernel_ab.py one cfg3 20 >> gpurun_out/${tag}_kernel_ab_deep.jsonl 2>&1
done
cut -c1-130 gpurun_out/${tag}_kernel_ab_deep.jsonl
timeout 60 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/${tag}_bench_n1.json 2>/dev/null | head -3
timeout 50 python tests/perf/fuzz_parity.py 50 9753 > gpurun_out/${tag}_fuzz.log 2>&1; tail -1 gpurun_out/${tag}_fuzz.log | cut -c1-250
