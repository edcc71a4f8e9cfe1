#!/bin/bash
# usage: tools/ncu_full.sh <tag> <path: single|pair|auto> <shape> [<shape> ...] — one `ncu --set full` capture (4th launch) per gemm_bench shape
tag=$1; path=$2; shift; shift
mkdir -p gpurun_out
for shape in "$@"; do
  timeout 120 python tools/gemm_bench.py --only $shape --iters 1 --path $path > gpurun_out/full_${tag}_$shape.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|gemm_pair' -s 3 -c 1 -f -o gpurun_out/full_${tag}_$shape \
      python tools/gemm_bench.py --only $shape --iters 1 --path $path >> gpurun_out/full_${tag}_$shape.log 2>&1
done
echo done
