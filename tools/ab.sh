#!/bin/bash
# usage: tools/ab.sh <tag> "<ENV=.. ENV=..>" ... — bench.py (no cpu baseline, no profile) under each env set; prints ms per step
tag=$1; shift
mkdir -p gpurun_out
: > gpurun_out/ab_$tag.log
for envs in "$@"; do
  out=$(env $envs timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>>gpurun_out/ab_$tag.err | tail -1)
  echo "$envs :: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("img/s %.2f  ms/step %.3f  e2e %.2f" % (d["value"], d["ms_per_unet_controlnet_step"], d["e2e"]["value"]))' 2>&1)" >> gpurun_out/ab_$tag.log
done
