#!/bin/bash
# usage: tools/gpu_round2.sh <tag> — gpu_round.sh plus the other BASELINE.json configs at their per-GPU shapes on one GPU
tag=${1:-x}
bash tools/gpu_round.sh $tag
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_$tag.log 2>&1
bash tools/bench_configs.sh 1 ${tag}_c2_b64 --batch 64 --no-cpu-baseline
bash tools/bench_configs.sh 1 ${tag}_c2_b128 --batch 128 --no-cpu-baseline
bash tools/bench_configs.sh 1 ${tag}_c3_b4 --size 512 --cfg 9 --batch 4 --no-cpu-baseline
bash tools/bench_configs.sh 1 ${tag}_c4_sweep --sweep --no-cpu-baseline
echo done2
