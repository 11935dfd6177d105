#!/bin/bash
# Round 2, trip 6: forward vector taps through the transposed copy, backward-MLP loads hoisted: tests + times + bench lines.
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
python scripts/kernel_times.py --segments 50 > gpurun_out/kernel_times6_50.txt 2>&1; grep -v touched gpurun_out/kernel_times6_50.txt
OUT=gpurun_out/trip6_ab.txt; : > $OUT
for m in train render image; do
  timeout 200 python bench.py --mode $m --steps 20 --warmup 5 --no-cpu-baseline --no-companions > gpurun_out/bench6_$m.json 2> gpurun_out/bench6_$m.err
  python -c "
import json
l=json.loads(open('gpurun_out/bench6_$m.json').readline()); print('$m', round(l['value'],1), l['unit'], round(l['ms_per_step'],4), 'ms', 'e2e', round(l['e2e']['value'],1), l.get('phases_ms'))" | tee -a $OUT
done
tail -n 3 gpurun_out/bench6_*.err
