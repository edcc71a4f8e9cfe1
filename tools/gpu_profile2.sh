#!/bin/bash
# usage: tools/gpu_profile2.sh <tag> — ncu launch list of one warm evaluation + full captures of the round's dominant kernels
tag=${1:-x}
mkdir -p gpurun_out
bash tools/ncu_list.sh $tag
bash tools/ncu_full.sh $tag auto conv320_c1 conv640_c2 sq320_plain conv1280_4x4
for k in gn_apply_stats; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/full_${tag}_$k \
      python tools/one_eval.py --evals 3 > gpurun_out/full_${tag}_$k.log 2>&1
done
echo done
