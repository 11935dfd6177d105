#!/bin/bash
# Data-parallel trip (run with gpurun --gpus N, N = 2 or 8): NCCL/P2P parity tests (2 ranks) + train bench at 1 and N GPUs
# for both exchange modes.  usage: trip_dp.sh <N>
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/dp_gpus.txt 2>&1
nvidia-smi topo -m > gpurun_out/dp_topo.txt 2>&1
timeout 900 python -m pytest tests/test_dp_gpu.py -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/test_dp_gpu_${N}gpu.log 2>&1
echo "test_dp_gpu exit=$? $(tail -1 gpurun_out/test_dp_gpu_${N}gpu.log)" | tee gpurun_out/summary_dp.txt
OUT=gpurun_out/dp_ab.txt; : > $OUT
run() { local n=$1 label=$2; shift 2
  if [ $n -eq 1 ]; then env "$@" timeout 200 python bench.py --mode train --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-companions > gpurun_out/dp_${label}.json 2> gpurun_out/dp_${label}.err
  else env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --mode train --gpus $n --steps 20 --warmup 5 > gpurun_out/dp_${label}.json 2> gpurun_out/dp_${label}.err; fi
  python -c "
import json
l=json.loads([x for x in open('gpurun_out/dp_${label}.json') if x.startswith('{')][-1]); print('$label', 'n', l['n_gpus'], 'ms_per_step', round(l['ms_per_step'],4), 'rays/s', round(l['value']), 'e2e', round(l['e2e']['value']), 'phases', {k: round(v,3) for k,v in l['phases_ms'].items()})" | tee -a $OUT || tail -5 gpurun_out/dp_${label}.err; }
run 1 n1
run $N p2p_n$N HRF_TRAIN_EXCHANGE=p2p
run $N nccl_n$N HRF_TRAIN_EXCHANGE=nccl
if [ $N -ge 4 ]; then run 2 p2p_n2 HRF_TRAIN_EXCHANGE=p2p; run 4 p2p_n4 HRF_TRAIN_EXCHANGE=p2p; fi
if [ $N -ge 8 ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29901 bench.py --mode sweep --gpus $N --steps 4 --warmup 3 > gpurun_out/dp_sweep_n$N.json 2> gpurun_out/dp_sweep_n$N.err; tail -c 900 gpurun_out/dp_sweep_n$N.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29902 bench.py --mode render --gpus $N --steps 20 --warmup 5 > gpurun_out/dp_render_n$N.json 2> gpurun_out/dp_render_n$N.err; tail -c 600 gpurun_out/dp_render_n$N.json
fi
tail -n 4 gpurun_out/dp_*.err | tail -40
