#!/bin/bash
# Round 2, trip 2: the sync-free training step (feature reuse, fused loss, multi-tensor Adam): tests + A/B of the reuse modes.
mkdir -p gpurun_out
for f in tests/test_training_gpu.py tests/test_composite_gpu.py tests/test_backward_gpu.py tests/test_scatter_gpu.py tests/test_field_gpu.py tests/test_integration_gpu.py tests/test_fullsize_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$? $(tail -1 gpurun_out/$n.log)" | tee -a gpurun_out/summary2.txt
done
OUT=gpurun_out/trip2_ab.txt; : > $OUT
train() { local label=$1; shift; env "$@" timeout 120 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/train_$label.err | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('train $label', 'ms_per_step', round(l['ms_per_step'],4), 'rays/s', round(l['value']), 'e2e', round(l['e2e']['value']), 'bwd_kernels_ms', round(l['roofline']['kernel_ms'],4), 'kept', l['config'].get('samples_after_prune_mean'))" | tee -a $OUT; }
train none HRF_TRAIN_REUSE=none
train feat HRF_TRAIN_REUSE=feat
train featgrid HRF_TRAIN_REUSE=feat+grid
tail -3 gpurun_out/train_*.err
