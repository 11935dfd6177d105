#!/bin/bash
# Round 2, first GPU trip: parity at BASELINE sizes + the whole GPU suite, per-kernel times, A/Bs prepared in round 1,
# ncu evidence of what ships at HEAD.
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
python scripts/kernel_times.py --segments 50 > gpurun_out/kernel_times_50.txt 2>&1; tail -25 gpurun_out/kernel_times_50.txt
python scripts/kernel_times.py --segments 100 100 50 --reps 5 > gpurun_out/kernel_times_250.txt 2>&1; tail -25 gpurun_out/kernel_times_250.txt
OUT=gpurun_out/round2_ab.txt; : > $OUT
bench() { local label=$1; shift; env "$@" timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-companions 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('render $label', 'kernel_ms', round(l['roofline']['kernel_ms'],4), 'ms_per_step', round(l['ms_per_step'],4), 'e2e', round(l['e2e']['value']))" | tee -a $OUT; }
train() { local label=$1; shift; env "$@" timeout 120 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('train $label', 'ms_per_step', round(l['ms_per_step'],4), 'bwd_kernels_ms', round(l['roofline']['kernel_ms'],4))" | tee -a $OUT; }
bench default HRF_FWD_UNROLL=1
bench ctas6 HRF_FWD_UNROLL=1 HRF_FWD_CTAS=6
train default HRF_BWD_UNROLL=0
train bwd_unroll1 HRF_BWD_UNROLL=1
bash scripts/gpu_profile.sh r2a
