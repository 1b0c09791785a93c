#!/bin/bash
# Same-box A/B of the deep-level kernel's grid (P2S_DEEP_GRID_MULT builds under pose2sim_b200/ab/, tools/kernel_ab.py build)
# and of the parking threshold, on the cfg3 shard, after the full GPU suite and the default bench line on the default build.
#   bash tools/deep_grid_ab.sh <tag>
tag=${1:-deep}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu -x > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/${tag}_pytest_gpu.log
timeout 120 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
timeout 150 python tools/kernel_ab.py run cfg3 20 > gpurun_out/${tag}_kernel_ab_deep_grid.jsonl 2>&1
for dm in 1000 600; do
    P2S_DEEP_MIN=$dm timeout 60 python tools/kernel_ab.py one cfg3 20 >> gpurun_out/${tag}_kernel_ab_deep_grid.jsonl 2>&1
done
cut -c1-130 gpurun_out/${tag}_kernel_ab_deep_grid.jsonl
python tools/show_bench.py gpurun_out/${tag}_bench_n1.json 2>/dev/null | head -5
