#!/bin/bash
# usage: tools/gpu_quick.sh <tag> [bench args] — kernel tests (under timeout), model tests, bench; no ncu
tag=${1:-x}
shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/k_$tag.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -s -rA 2>&1 | grep -E "rel-L2|PSNR|passed|failed|PASSED|FAILED|Error|assert" > gpurun_out/m_$tag.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/conv_table_$tag.json "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo done
