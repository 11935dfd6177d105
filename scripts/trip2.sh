#!/bin/bash
# Round 2, trip 2: the sync-free training step (feature reuse, fused loss, multi-tensor Adam) and the parity-slot scatter:
# tests + A/B of the reuse modes and the scatter generations.
mkdir -p gpurun_out
: > gpurun_out/summary2.txt
for f in tests/test_scatter_gpu.py tests/test_training_gpu.py tests/test_composite_gpu.py tests/test_backward_gpu.py tests/test_baseline_sizes_gpu.py tests/test_field_gpu.py tests/test_integration_gpu.py tests/test_fullsize_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$? $(tail -1 gpurun_out/$n.log)" | tee -a gpurun_out/summary2.txt
done
python scripts/kernel_times.py --segments 50 > gpurun_out/kernel_times2_50.txt 2>&1; grep -i "scatter\|prune pass" gpurun_out/kernel_times2_50.txt
HRF_SCATTER_CTAS=5 python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter" | sed 's/^/ctas5 /' | tee -a gpurun_out/kernel_times2_50.txt
HRF_SCATTER_V2=0 python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter" | sed 's/^/v1 /' | tee -a gpurun_out/kernel_times2_50.txt
OUT=gpurun_out/trip2_ab.txt; : > $OUT
train() { local label=$1; shift; env "$@" timeout 120 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/train_$label.err | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('train $label', 'ms_per_step', round(l['ms_per_step'],4), 'rays/s', round(l['value']), 'e2e', round(l['e2e']['value']), 'bwd_kernels_ms', round(l['roofline']['kernel_ms'],4), 'kept', l['config'].get('samples_after_prune_mean'))" | tee -a $OUT; }
train none HRF_TRAIN_REUSE=none
train feat HRF_TRAIN_REUSE=feat
train featgrid HRF_TRAIN_REUSE=feat+grid
train featgrid_ctas5 HRF_TRAIN_REUSE=feat+grid HRF_SCATTER_CTAS=5
train featgrid_v1 HRF_TRAIN_REUSE=feat+grid HRF_SCATTER_V2=0
tail -3 gpurun_out/train_*.err
timeout 300 python bench.py --mode image --steps 3 --warmup 3 > gpurun_out/b_image2.json 2> gpurun_out/b_image2.err; cat gpurun_out/b_image2.json; tail -3 gpurun_out/b_image2.err
timeout 200 python bench.py --mode render --steps 20 --warmup 5 --no-cpu-baseline --no-companions > gpurun_out/b_render2.json 2> gpurun_out/b_render2.err; cat gpurun_out/b_render2.json; tail -3 gpurun_out/b_render2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grid_scatter_v2" -s 8 -c 1 -o gpurun_out/prof_r2b_scatter -f \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_scatter2.log 2>&1
