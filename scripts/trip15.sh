#!/bin/bash
# Round 2: render-mode bench with int32 ray indices in the e2e upload.
mkdir -p gpurun_out
timeout 300 python bench.py --mode render --no-cpu-baseline --no-companions > gpurun_out/r2o_bench_render.json 2> gpurun_out/r2o_bench_render.err
python -c "
import json
l=json.loads(open('gpurun_out/r2o_bench_render.json').readline()); print('render', round(l['value'],1), round(l['ms_per_step'],4), 'ms', 'e2e', l['e2e'])" | tee gpurun_out/summary15.txt
tail -3 gpurun_out/r2o_bench_render.err
