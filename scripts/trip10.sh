#!/bin/bash
# Round 2, trip 10: scatter v5 + warp-level sum of the time-axis vector gradient; batch (frame pool) sensitivity.
mkdir -p gpurun_out
: > gpurun_out/summary10.txt
timeout 600 python -m pytest tests/test_scatter_gpu.py -q -m gpu -s -x --no-header -p no:cacheprovider -k "v5 or v3-5ctas" > gpurun_out/test_scatter_gpu10.log 2>&1
echo "test_scatter_gpu exit=$? $(tail -1 gpurun_out/test_scatter_gpu10.log)" | tee -a gpurun_out/summary10.txt
for t in tests/test_backward_gpu.py tests/test_training_gpu.py; do
  n=$(basename $t .py)
  HRF_SCATTER=5 timeout 900 python -m pytest $t -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/${n}_v5b.log 2>&1
  echo "${n}_v5 exit=$? $(tail -1 gpurun_out/${n}_v5b.log)" | tee -a gpurun_out/summary10.txt
done
HRF_SCATTER=5 timeout 300 python scripts/seed_spread.py 2>&1 | tee gpurun_out/seed_spread_v5.txt | tail -8
HRF_SCATTER=5 python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter" | sed "s/^/gen5b /" | tee gpurun_out/kernel_times10_50.txt
