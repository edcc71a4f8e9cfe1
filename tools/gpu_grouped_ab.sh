#!/bin/bash
# usage: tools/gpu_grouped_ab.sh <tag> — model tests (the grouped trunk is the default path), then bench A/B: stacked trunk vs two networks
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -s -rA 2>&1 | grep -E "rel-L2|PSNR|passed|failed|PASSED|FAILED|Error|assert" > gpurun_out/m_$tag.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/conv_table_$tag.json > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-grouped > gpurun_out/bench_${tag}_nogroup.json 2> gpurun_out/bench_${tag}_nogroup.err
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/k_$tag.log
tail -3 gpurun_out/m_$tag.log; cat gpurun_out/bench_$tag.json | head -c 600; echo; cat gpurun_out/bench_${tag}_nogroup.json | head -c 600; echo; tail -3 gpurun_out/bench_$tag.err
echo done
