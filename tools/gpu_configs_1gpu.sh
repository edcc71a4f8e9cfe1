#!/bin/bash
# usage: tools/gpu_configs_1gpu.sh <tag> — the other BASELINE configs' per-GPU shapes on one B200, stacked trunk (auto / forced) vs two networks
tag=${1:-x}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_${tag}_$name.json 2> gpurun_out/bench_${tag}_$name.err; }
run b64 --batch 64
run b64_sep --batch 64 --no-grouped
run c3 --size 512 --cfg 9 --batch 4
run c3_sep --size 512 --cfg 9 --batch 4 --no-grouped
run sweep --sweep
run sweep_forced --sweep --grouped
python - "$tag" <<'PY'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/bench_%s_*.json" % sys.argv[1])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(d["value"], 2), round(d["ms_per_unet_controlnet_step"], 3), round(d["roofline"]["frac"], 3), d["trunks"][:40])
    except Exception as e:
        print(f, "ERR", e)
PY
echo done
