#!/bin/bash
# Round 2, trip 7: scatter v3 third cut (segment descriptors in shared memory, df double-buffered with cp.async).
mkdir -p gpurun_out
: > gpurun_out/summary7.txt
for f in tests/test_scatter_gpu.py tests/test_training_gpu.py tests/test_backward_gpu.py tests/test_baseline_sizes_gpu.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n exit=$? $(tail -1 gpurun_out/$n.log)" | tee -a gpurun_out/summary7.txt
done
python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter" | tee gpurun_out/kernel_times7_50.txt
timeout 200 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline --no-companions > gpurun_out/bench7_train.json 2> gpurun_out/bench7_train.err
python -c "
import json
l=json.loads(open('gpurun_out/bench7_train.json').readline()); print('train', round(l['value'],1), l['unit'], round(l['ms_per_step'],4), 'ms', 'e2e', round(l['e2e']['value'],1), l.get('phases_ms'))" | tee gpurun_out/trip7_ab.txt
tail -n 3 gpurun_out/bench7_train.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"grid_scatter_v3" -s 8 -c 1 -o gpurun_out/prof_r2g_scatter -f \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_full_scatter7.log 2>&1
