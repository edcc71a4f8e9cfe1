#!/bin/bash
# timing-only experiments on the GEMM epilogue (RESULTS INVALID): which part of a launch is loads / stores / MMA
# needs a library built with the switches compiled in: MKD_TRACE=1 python -m makeupdiffuse_b200.build --force
for v in 0 4 8 12 2 14; do
  echo "== MKD_DEBUG_TIMING=$v  (1 skip B loads, 2 skip MMA issue, 4 skip epilogue stores, 8 skip residual loads)"
  MKD_DEBUG_TIMING=$v timeout 300 python tools/gemm_bench.py 2>&1 | grep -E "sq320|qkv|conv320|ff2|sq1280_res|sq640"
done
