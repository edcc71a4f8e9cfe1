#!/bin/bash
# usage: tools/gpu_evidence.sh <tag> — the round's single-GPU evidence set: full GPU test suite, smoke, default bench (with the
# CPU baseline), the reference arm, kernel micro-benchmarks; everything into gpurun_out/
tag=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -rA 2>&1 | grep -E "rel-L2|PSNR|passed|failed|PASSED|FAILED|SKIPPED|Error|assert" > gpurun_out/gputests_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1
timeout 900 python bench.py --profile-out gpurun_out/conv_table_$tag.json > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_$tag.log 2>&1
timeout 300 python tools/norm_bench.py > gpurun_out/norm_$tag.log 2>&1
timeout 300 python tools/attn_bench.py > gpurun_out/attn_$tag.log 2>&1
echo done
