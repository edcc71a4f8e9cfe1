#!/bin/bash
# usage: tools/gpu_small_batch_ab.sh <tag> — per-GPU shapes of configs[4] on 8 / 4 / 2 GPUs (5 / 10 / 20 samples, DDIM-20): "auto" vs stacked forced
tag=${1:-x}
mkdir -p gpurun_out
for b in 5 10 20; do
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch $b --ddim-steps 20 > gpurun_out/bench_${tag}_b${b}_auto.json 2> gpurun_out/bench_${tag}_b${b}_auto.err
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch $b --ddim-steps 20 --grouped > gpurun_out/bench_${tag}_b${b}_forced.json 2> gpurun_out/bench_${tag}_b${b}_forced.err
done
python - "$tag" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/bench_%s_*.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d["roofline"]
        print(f.split("/")[-1], round(d["value"], 2), round(d["ms_per_unet_controlnet_step"], 3), round(r["frac"], 4), d["gpu_launches"], d["trunks"][:30])
    except Exception as e:
        print(f, "ERR", e)
PY
