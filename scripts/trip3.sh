#!/bin/bash
# Round 2, trip 3: scatter v2 (second cut) A/B, cold vs warm L2, train step variants with full JSON lines.
mkdir -p gpurun_out
: > gpurun_out/summary3.txt
for f in tests/test_scatter_gpu.py tests/test_training_gpu.py tests/test_composite_gpu.py tests/test_integration_gpu.py tests/test_backward_gpu.py tests/test_baseline_sizes_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$? $(tail -1 gpurun_out/$n.log)" | tee -a gpurun_out/summary3.txt
done
python scripts/kernel_times.py --segments 50 > gpurun_out/kernel_times3_50.txt 2>&1; cat gpurun_out/kernel_times3_50.txt | grep -v touched
HRF_SCATTER_CTAS=5 python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter" | sed 's/^/ctas5 /' | tee -a gpurun_out/kernel_times3_50.txt
HRF_SCATTER_V2=0 python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter" | sed 's/^/v1 /' | tee -a gpurun_out/kernel_times3_50.txt
OUT=gpurun_out/trip3_ab.txt; : > $OUT
train() { local label=$1; shift; env "$@" timeout 150 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline --no-companions 2>gpurun_out/train3_$label.err > gpurun_out/train3_$label.json; python -c "
import json,sys
l=json.loads(open('gpurun_out/train3_$label.json').readline()); print('train $label', 'ms_per_step', round(l['ms_per_step'],4), 'rays/s', round(l['value']), 'e2e', round(l['e2e']['value']), 'phases', {k: round(v,3) for k,v in l['phases_ms'].items()})" | tee -a $OUT; }
train featgrid HRF_TRAIN_REUSE=feat+grid
train featgrid_ctas5 HRF_TRAIN_REUSE=feat+grid HRF_SCATTER_CTAS=5
train featgrid_v1 HRF_TRAIN_REUSE=feat+grid HRF_SCATTER_V2=0
train feat_ctas5 HRF_TRAIN_REUSE=feat HRF_SCATTER_CTAS=5
tail -n 3 gpurun_out/train3_*.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grid_scatter_v2" -s 8 -c 1 -o gpurun_out/prof_r2c_scatter -f \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_full_scatter3.log 2>&1
