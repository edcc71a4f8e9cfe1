#!/bin/bash
# usage: tools/gpu_round.sh <tag>   — tests, bench, and an ncu launch list of one eval, all into gpurun_out/
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -6 > gpurun_out/k_$tag.log
timeout 900 python -m pytest tests/test_model_gpu.py -q -s -rA 2>&1 | grep -E "rel-L2|PSNR|passed|failed|PASSED|FAILED|Error|assert" > gpurun_out/m_$tag.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/conv_table_$tag.json > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
timeout 300 python tools/one_eval.py --evals 2 > gpurun_out/one_eval_$tag.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/one_eval.py --evals 2 > gpurun_out/ncu_$tag.log 2>&1
echo done
