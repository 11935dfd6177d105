#!/bin/bash
# Round 2: int32 ray indices in the e2e upload (FusedTrainer.step widens them): test + train bench line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_training_gpu.py -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/test_training_gpu_14.log 2>&1
echo "test_training_gpu exit=$? $(tail -1 gpurun_out/test_training_gpu_14.log)" | tee gpurun_out/summary14.txt
timeout 300 python bench.py --mode train --no-cpu-baseline --no-companions > gpurun_out/r2o_bench_train.json 2> gpurun_out/r2o_bench_train.err
python -c "
import json
l=json.loads(open('gpurun_out/r2o_bench_train.json').readline()); print('train', round(l['value'],1), round(l['ms_per_step'],4), 'ms', 'e2e', l['e2e'])" | tee -a gpurun_out/summary14.txt
tail -3 gpurun_out/r2o_bench_train.err
