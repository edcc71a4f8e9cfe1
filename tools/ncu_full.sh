#!/bin/bash
# usage: tools/ncu_full.sh <tag> <shape> [<shape> ...] — one `ncu --set full` capture (4th launch) per gemm_bench shape
tag=$1; shift
mkdir -p gpurun_out
for shape in "$@"; do
  timeout 120 python tools/gemm_bench.py --only $shape --iters 1 > gpurun_out/full_${tag}_$shape.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -f -o gpurun_out/full_${tag}_$shape \
      python tools/gemm_bench.py --only $shape --iters 1 >> gpurun_out/full_${tag}_$shape.log 2>&1
done
echo done
