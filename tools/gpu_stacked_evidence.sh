#!/bin/bash
# usage: tools/gpu_stacked_evidence.sh <tag> — GPU test suite, gemm_bench of the stacked (wgroups = 2) launches, full ncu captures of two of them
tag=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -rA 2>&1 | grep -E "rel-L2|PSNR|passed|failed|PASSED|FAILED|SKIPPED|Error|assert" > gpurun_out/gputests_$tag.log
timeout 300 python tools/gemm_bench.py --only g2_ > gpurun_out/gemm_g2_$tag.log 2>&1
bash tools/ncu_full.sh $tag auto g2_conv320_c1 g2_sq320_res32 g2_conv1280_4x4
for s in g2_conv320_c1 g2_sq320_res32 g2_conv1280_4x4; do
  [ -f gpurun_out/full_${tag}_$s.ncu-rep ] && bash tools/ncu_rep_summary.sh gpurun_out/full_${tag}_$s.ncu-rep "gemm_pair_kernel, stacked launch $s (wgroups = 2)" > gpurun_out/full_${tag}_$s.txt
done
tail -2 gpurun_out/gputests_$tag.log; cat gpurun_out/gemm_g2_$tag.log | tail -12
echo done
