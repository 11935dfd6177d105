#!/bin/bash
# Round 2, trip 9: scatter v5 (lane pairs, pair-wide flushes -> x-neighbour entries share one L2 request) vs v3; per-rank batch spread.
mkdir -p gpurun_out
: > gpurun_out/summary9.txt
for f in tests/test_scatter_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$? $(tail -1 gpurun_out/$n.log)" | tee -a gpurun_out/summary9.txt
done
for t in tests/test_backward_gpu.py tests/test_baseline_sizes_gpu.py tests/test_training_gpu.py; do
  n=$(basename $t .py)
  HRF_SCATTER=5 timeout 900 python -m pytest $t -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/${n}_v5.log 2>&1
  echo "${n}_v5 exit=$? $(tail -1 gpurun_out/${n}_v5.log)" | tee -a gpurun_out/summary9.txt
done
: > gpurun_out/kernel_times9_50.txt
for g in 3 5; do HRF_SCATTER=$g python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter" | sed "s/^/gen$g /" | tee -a gpurun_out/kernel_times9_50.txt; done
OUT=gpurun_out/trip9_ab.txt; : > $OUT
for g in 3 5; do
  HRF_SCATTER=$g timeout 200 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline --no-companions > gpurun_out/bench9_train_g$g.json 2> gpurun_out/bench9_train_g$g.err
  python -c "
import json
l=json.loads(open('gpurun_out/bench9_train_g$g.json').readline()); print('train gen$g', round(l['value'],1), l['unit'], round(l['ms_per_step'],4), 'ms', 'e2e', round(l['e2e']['value'],1), l.get('phases_ms'))" | tee -a $OUT
done
tail -n 3 gpurun_out/bench9_train_g*.err
timeout 300 python scripts/seed_spread.py 2>&1 | tee gpurun_out/seed_spread.txt | tail -8
HRF_SCATTER=5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grid_scatter_v5" -s 8 -c 1 -o gpurun_out/prof_r2j_scatter -f \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_full_scatter9.log 2>&1
tail -2 gpurun_out/ncu_full_scatter9.log | cut -c1-200
