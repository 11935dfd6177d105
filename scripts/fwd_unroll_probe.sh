#!/bin/bash
# A/B of the encode loop's level unroll (HRF_FWD_UNROLL = 4 | 2 | 1) on the render bench, then the field parity tests
# with the fastest setting exported.
mkdir -p gpurun_out
: > gpurun_out/fwd_unroll.txt
for u in 4 2 1; do
  HRF_FWD_UNROLL=$u timeout 40 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('unroll $u', l['roofline']['kernel_ms'], l['ms_per_step'])" | tee -a gpurun_out/fwd_unroll.txt
done
BEST=$(sort -k3 -n gpurun_out/fwd_unroll.txt | head -1 | awk '{print $2}')
echo "best $BEST" | tee -a gpurun_out/fwd_unroll.txt
HRF_FWD_UNROLL=$BEST timeout 60 python -m pytest tests/test_field_gpu.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2 | tee -a gpurun_out/fwd_unroll.txt
