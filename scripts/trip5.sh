#!/bin/bash
# Round 2, trip 5: scatter v3 second cut (egrid columns fetched once, vector taps prefetched one step ahead).
mkdir -p gpurun_out
: > gpurun_out/summary5.txt
for f in tests/test_scatter_gpu.py tests/test_training_gpu.py tests/test_backward_gpu.py tests/test_baseline_sizes_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$? $(tail -1 gpurun_out/$n.log)" | tee -a gpurun_out/summary5.txt
done
: > gpurun_out/kernel_times5_50.txt
for c in 5 6; do HRF_SCATTER=3 HRF_SCATTER_CTAS=$c python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter\|backward MLP\|prune pass +" | sed "s/^/gen3 ctas$c /" | tee -a gpurun_out/kernel_times5_50.txt; done
OUT=gpurun_out/trip5_ab.txt; : > $OUT
train() { local label=$1; shift; env "$@" timeout 150 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline --no-companions 2>gpurun_out/train5_$label.err > gpurun_out/train5_$label.json; python -c "
import json,sys
l=json.loads(open('gpurun_out/train5_$label.json').readline()); print('train $label', 'ms_per_step', round(l['ms_per_step'],4), 'rays/s', round(l['value']), 'e2e', round(l['e2e']['value']), 'phases', {k: round(v,3) for k,v in l['phases_ms'].items()})" | tee -a $OUT; }
train default
train ctas6 HRF_SCATTER_CTAS=6
tail -n 3 gpurun_out/train5_*.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grid_scatter_v3" -s 8 -c 1 -o gpurun_out/prof_r2e_scatter -f \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_full_scatter5.log 2>&1
