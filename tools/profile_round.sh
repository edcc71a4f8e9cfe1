#!/bin/bash
# usage: tools/profile_round.sh <tag> — the round's evidence: ncu launch list (+ DRAM bytes) of one warm evaluation and
# `ncu --set full` captures of the dominant kernels (each only after the same command ran clean without ncu)
tag=${1:-x}
mkdir -p gpurun_out
bash tools/ncu_list.sh $tag
bash tools/ncu_full.sh $tag conv320 ff1_320 sq320_res32 conv1280
timeout 120 python tools/attn_bench.py "self 32x32 d40" > gpurun_out/full_${tag}_attn.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tcgen05 -s 3 -c 1 -f -o gpurun_out/full_${tag}_attn \
    python tools/attn_bench.py "self 32x32 d40" >> gpurun_out/full_${tag}_attn.log 2>&1
for k in gn_apply_stats layernorm5 gn_group; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:$k -c 1 -f -o gpurun_out/full_${tag}_$k \
      python tools/one_eval.py --evals 3 > gpurun_out/full_${tag}_$k.log 2>&1
done
echo done
