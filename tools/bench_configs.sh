#!/bin/bash
# usage: tools/bench_configs.sh <N> <tag> [bench args]  — one bench.py run on N GPUs of this box, JSON line into gpurun_out/bench_<tag>.json
N=$1; tag=$2; shift; shift
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 --steps 3 --warmup 3 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
fi
tail -c 300 gpurun_out/bench_$tag.err
