#!/bin/bash
# First GPU trip of the next round: re-measure what changed after the last GPU minute of round 1 and run the prepared
# A/Bs.  One B200, ~4 min.  Results under gpurun_out/round2_ab.txt.
#   gpurun --timeout 600 -- 'bash scripts/round2_experiments.sh'
mkdir -p gpurun_out
OUT=gpurun_out/round2_ab.txt
: > $OUT
bench() {  # label, env assignments...
  local label=$1; shift
  env "$@" timeout 90 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-companions 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('render $label', 'kernel_ms', round(l['roofline']['kernel_ms'],4), 'ms_per_step', round(l['ms_per_step'],4), 'e2e', round(l['e2e']['value']))" | tee -a $OUT
}
train() {
  local label=$1; shift
  env "$@" timeout 90 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('train $label', 'ms_per_step', round(l['ms_per_step'],4), 'bwd_kernels_ms', round(l['roofline']['kernel_ms'],4))" | tee -a $OUT
}
# 1. parity first (the index-wrap loop and the multi-segment optimiser path were committed after the last full GPU run)
timeout 300 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2 | tee -a $OUT
# 2. forward: default (1-level loop + wrap loop), the old 4-level loop, 6 CTAs/SM
bench default HRF_FWD_UNROLL=1
bench unroll4 HRF_FWD_UNROLL=4
bench ctas6 HRF_FWD_UNROLL=1 HRF_FWD_CTAS=6
# 3. training step: default, compact backward kernel
train default HRF_BWD_UNROLL=0
train bwd_unroll1 HRF_BWD_UNROLL=1
HRF_BWD_UNROLL=1 timeout 120 python -m pytest tests/test_backward_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -1 | tee -a $OUT
