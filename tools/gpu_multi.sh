#!/bin/bash
# usage: tools/gpu_multi.sh <N> <tag> [configs...] — BASELINE.json configs[1..4] on N GPUs of this box (default: all), JSON lines into gpurun_out/
N=$1; tag=$2; shift; shift
cfgs=${@:-c1 c2 c3 c4}
for c in $cfgs; do
  case $c in
    c1) bash tools/bench_configs.sh $N ${tag}_c1_n$N --no-cpu-baseline ;;
    c2) bash tools/bench_configs.sh $N ${tag}_c2_n$N --global-batch 128 --no-cpu-baseline ;;
    c3) bash tools/bench_configs.sh $N ${tag}_c3_n$N --size 512 --cfg 9 --global-batch 32 --no-cpu-baseline ;;
    c4) bash tools/bench_configs.sh $N ${tag}_c4_n$N --sweep --no-cpu-baseline ;;
    t)  timeout 600 python -m pytest tests/test_sharding.py -q -m gpu 2>&1 | tail -3 > gpurun_out/sharding_${tag}_n$N.log ;;
  esac
done
echo done
