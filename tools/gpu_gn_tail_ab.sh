#!/bin/bash
# usage: tools/gpu_gn_tail_ab.sh <tag> — bench A/B on one box: GroupNorm tail on / off, twice each (interleaved)
tag=${1:-x}
mkdir -p gpurun_out
for i in 1 2; do
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_tail$i.json 2> gpurun_out/bench_${tag}_tail$i.err
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gn-tail > gpurun_out/bench_${tag}_notail$i.json 2> gpurun_out/bench_${tag}_notail$i.err
done
python - "$tag" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/bench_%s_*.json" % sys.argv[1])):
    d = json.loads(open(f).read().strip().splitlines()[-1]); r = d["roofline"]
    print(f.split("/")[-1], round(d["value"], 2), round(d["ms_per_unet_controlnet_step"], 3), round(r["frac"], 4), r["launches_per_eval"], round(r["kernel_ms_per_eval"], 3),
          d["gpu_launches"], r["others"]["groupnorm"]["launches_per_eval"], round(r["others"]["groupnorm"]["kernel_ms_per_eval"], 3))
PY
