#!/bin/bash
# Round 2, trip 12: vector-row gradients accumulated in the transposed scratch (v3) + fold.
mkdir -p gpurun_out
: > gpurun_out/summary12.txt
for t in tests/test_scatter_gpu.py tests/test_training_gpu.py tests/test_backward_gpu.py; do
  n=$(basename $t .py)
  timeout 900 python -m pytest $t -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/${n}_12.log 2>&1
  echo "${n} exit=$? $(tail -1 gpurun_out/${n}_12.log)" | tee -a gpurun_out/summary12.txt
done
python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter" | tee gpurun_out/kernel_times12_50.txt
OUT=gpurun_out/trip12_ab.txt; : > $OUT
for v in 0 1; do
  HRF_VECGRAD_T=$v timeout 200 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline --no-companions > gpurun_out/bench12_train_vt$v.json 2> gpurun_out/bench12_train_vt$v.err
  python -c "
import json
l=json.loads(open('gpurun_out/bench12_train_vt$v.json').readline()); print('train vecgrad_t=$v', round(l['value'],1), l['unit'], round(l['ms_per_step'],4), 'ms', 'e2e', round(l['e2e']['value'],1), l.get('phases_ms'))" | tee -a $OUT
done
tail -n 3 gpurun_out/bench12_train_vt*.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grid_scatter_v3" -s 8 -c 1 -o gpurun_out/prof_r2n_scatter_v3 -f \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_full_scatter12.log 2>&1
