#!/bin/bash
# usage: tools/gpu_plan_ab.sh <tag> — kernel tests, the stacked micro-benchmark, one bench line (plan-level A/B of the pair kernel)
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -3 > gpurun_out/k_$tag.log
timeout 300 python tools/gemm_bench.py --only g2_ > gpurun_out/gemm_g2_$tag.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cat gpurun_out/k_$tag.log; cat gpurun_out/gemm_g2_$tag.log; head -c 250 gpurun_out/bench_$tag.json; echo
