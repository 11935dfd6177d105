#!/bin/bash
# Runs every GPU test file in its own process (a sticky CUDA error in one file must not poison the rest),
# logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_umma_selftest_gpu.py tests/test_texture_probe_gpu.py tests/test_sampler_gpu.py tests/test_composite_gpu.py \
         tests/test_field_gpu.py tests/test_backward_gpu.py tests/test_ref_parity_gpu.py tests/test_training_gpu.py tests/test_data_loader_gpu.py tests/test_integration_gpu.py tests/test_fullsize_gpu.py tests/test_dp_gpu.py tests/test_occupancy_tools_gpu.py tests/test_scatter_gpu.py; do
  [ -f "$f" ] || continue
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$?" | tee -a gpurun_out/summary.txt
  tail -3 gpurun_out/$n.log
done
