#!/bin/bash
# 2-GPU trip: head-of-step all-reduce made asynchronous (runs beside the prune pass): DP parity + train bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/test_dp_gpu_2gpu_d.log 2>&1
echo "test_dp_gpu exit=$? $(tail -1 gpurun_out/test_dp_gpu_2gpu_d.log)" | tee gpurun_out/summary_dp2d.txt
OUT=gpurun_out/dp2d_ab.txt; : > $OUT
run() { local n=$1 label=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --mode train --gpus $n --steps 20 --warmup 5 > gpurun_out/dp2d_${label}.json 2> gpurun_out/dp2d_${label}.err
  python -c "
import json
l=json.loads([x for x in open('gpurun_out/dp2d_${label}.json') if x.startswith('{')][-1]); print('$label', 'n', l['n_gpus'], 'ms_per_step', round(l['ms_per_step'],4), 'rays/s', round(l['value']), 'e2e', round(l['e2e']['value']), 'phases', {k: round(v,3) for k,v in l['phases_ms'].items()})" | tee -a $OUT || tail -5 gpurun_out/dp2d_${label}.err; }
run 2 p2p_serial_n2 HRF_TRAIN_EXCHANGE=p2p
run 2 nccl_n2 HRF_TRAIN_EXCHANGE=nccl
tail -n 4 gpurun_out/dp2d_*.err | tail -20
