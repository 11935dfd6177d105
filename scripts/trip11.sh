#!/bin/bash
# Round 2, trip 11: time-axis vector gradient summed across the warp in v3 / v4 / v5; 16-byte REDs in the MLP weight-gradient flush.
mkdir -p gpurun_out
: > gpurun_out/summary11.txt
timeout 900 python -m pytest tests/test_scatter_gpu.py -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/test_scatter_gpu11.log 2>&1
echo "test_scatter_gpu exit=$? $(tail -1 gpurun_out/test_scatter_gpu11.log)" | tee -a gpurun_out/summary11.txt
for g in 3 5; do
for t in tests/test_backward_gpu.py tests/test_training_gpu.py tests/test_baseline_sizes_gpu.py; do
  n=$(basename $t .py)
  HRF_SCATTER=$g timeout 900 python -m pytest $t -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/${n}_g$g.log 2>&1
  echo "${n}_g$g exit=$? $(tail -1 gpurun_out/${n}_g$g.log)" | tee -a gpurun_out/summary11.txt
done
done
: > gpurun_out/kernel_times11_50.txt
for g in 3 4 5; do HRF_SCATTER=$g python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter\|backward MLP" | sed "s/^/gen$g /" | tee -a gpurun_out/kernel_times11_50.txt; done
OUT=gpurun_out/trip11_ab.txt; : > $OUT
for g in 3 5; do
  HRF_SCATTER=$g timeout 200 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline --no-companions > gpurun_out/bench11_train_g$g.json 2> gpurun_out/bench11_train_g$g.err
  python -c "
import json
l=json.loads(open('gpurun_out/bench11_train_g$g.json').readline()); print('train gen$g', round(l['value'],1), l['unit'], round(l['ms_per_step'],4), 'ms', 'e2e', round(l['e2e']['value'],1), l.get('phases_ms'))" | tee -a $OUT
done
tail -n 3 gpurun_out/bench11_train_g*.err
for g in 3 5; do
HRF_SCATTER=$g timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grid_scatter_v$g" -s 8 -c 1 -o gpurun_out/prof_r2k_scatter_v$g -f \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_full_scatter11_v$g.log 2>&1
done
