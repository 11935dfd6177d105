#!/bin/bash
# Round 2, trip 4: scatter generations A/B (v1 staged / v2 / v3), e2e loop check, whole default bench line.
mkdir -p gpurun_out
: > gpurun_out/summary4.txt
for f in tests/test_scatter_gpu.py tests/test_training_gpu.py tests/test_backward_gpu.py tests/test_dp_gpu.py tests/test_baseline_sizes_gpu.py tests/test_field_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$? $(tail -1 gpurun_out/$n.log)" | tee -a gpurun_out/summary4.txt
done
: > gpurun_out/kernel_times4_50.txt
for g in 1 2 3; do for c in 6 5; do
  [ $g -eq 1 ] && [ $c -eq 5 ] && continue
  HRF_SCATTER=$g HRF_SCATTER_CTAS=$c python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter" | sed "s/^/gen$g ctas$c /" | tee -a gpurun_out/kernel_times4_50.txt
done; done
OUT=gpurun_out/trip4_ab.txt; : > $OUT
train() { local label=$1; shift; env "$@" timeout 150 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline --no-companions 2>gpurun_out/train4_$label.err > gpurun_out/train4_$label.json; python -c "
import json,sys
l=json.loads(open('gpurun_out/train4_$label.json').readline()); print('train $label', 'ms_per_step', round(l['ms_per_step'],4), 'rays/s', round(l['value']), 'e2e', round(l['e2e']['value']), 'phases', {k: round(v,3) for k,v in l['phases_ms'].items()})" | tee -a $OUT; }
train gen1 HRF_SCATTER=1
train gen3 HRF_SCATTER=3
train gen3_ctas5 HRF_SCATTER=3 HRF_SCATTER_CTAS=5
train gen1_omp1 HRF_SCATTER=1 OMP_NUM_THREADS=1
tail -n 3 gpurun_out/train4_*.err
timeout 600 python bench.py > gpurun_out/bench4_default.json 2> gpurun_out/bench4_default.err; tail -c 1500 gpurun_out/bench4_default.json; tail -3 gpurun_out/bench4_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grid_scatter_v3" -s 8 -c 1 -o gpurun_out/prof_r2d_scatter -f \
    env HRF_SCATTER=3 python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_full_scatter4.log 2>&1
