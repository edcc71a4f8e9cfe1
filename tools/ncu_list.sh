#!/bin/bash
# usage: tools/ncu_list.sh <tag> — ncu launch list (gpu__time_duration) of ONE warm UNet+ControlNet evaluation
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 python tools/one_eval.py --evals 3 > gpurun_out/one_eval_$tag.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$tag.csv python tools/one_eval.py --evals 3 > gpurun_out/ncu_$tag.log 2>&1
echo done
