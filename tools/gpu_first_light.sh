#!/bin/bash
# First GPU contact: staged so that a hanging/faulting tcgen05 kernel cannot take the box down with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== kernels (no tcgen05) ==" | tee gpurun_out/first_light.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "not tcgen05 and not auto_dispatch" 2>&1 | tail -25 | tee -a gpurun_out/first_light.log
for s in a b c d e f g h; do
  echo "== tc_probe $s ==" | tee -a gpurun_out/first_light.log
  timeout 90 python tools/tc_probe.py $s 2>&1 | tail -12 | tee -a gpurun_out/first_light.log
  echo "exit=$?" | tee -a gpurun_out/first_light.log
done
echo "== tcgen05 kernel tests ==" | tee -a gpurun_out/first_light.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "tcgen05 or auto_dispatch" 2>&1 | tail -40 | tee -a gpurun_out/first_light.log
echo "== model tests ==" | tee -a gpurun_out/first_light.log
timeout 1500 python -m pytest tests/test_model_gpu.py -q -s 2>&1 | tail -60 | tee -a gpurun_out/first_light.log
