#!/bin/bash
# 8-GPU trip (gpurun --gpus 8): train bench at 1 / 4 / 8 GPUs (P2P exchange), NCCL exchange at 8, the novel-view sweep and
# the fused render at 8.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/dp8_topo.txt 2>&1
OUT=gpurun_out/dp8_ab.txt; : > $OUT
run() { local n=$1 label=$2 mode=$3; shift 3
  if [ $n -eq 1 ]; then env "$@" timeout 200 python bench.py --mode $mode --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-companions > gpurun_out/dp8_${label}.json 2> gpurun_out/dp8_${label}.err
  else env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --mode $mode --gpus $n --steps 20 --warmup 5 > gpurun_out/dp8_${label}.json 2> gpurun_out/dp8_${label}.err; fi
  python -c "
import json
l=json.loads([x for x in open('gpurun_out/dp8_${label}.json') if x.startswith('{')][-1]); print('$label', 'n', l['n_gpus'], 'ms_per_step', round(l['ms_per_step'],4), l['unit'], round(l['value'],1), 'e2e', round(l['e2e']['value'],1), 'phases', {k: round(v,3) for k,v in (l.get('phases_ms') or {}).items()})" | tee -a $OUT || tail -5 gpurun_out/dp8_${label}.err; }
run 1 train_n1 train
run 8 train_p2p_n8 train HRF_TRAIN_EXCHANGE=p2p
run 8 train_nccl_n8 train HRF_TRAIN_EXCHANGE=nccl
run 4 train_p2p_n4 train HRF_TRAIN_EXCHANGE=p2p
run 8 render_n8 render
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29901 bench.py --mode sweep --gpus 8 --steps 4 --warmup 3 > gpurun_out/dp8_sweep_n8.json 2> gpurun_out/dp8_sweep_n8.err; tail -c 700 gpurun_out/dp8_sweep_n8.json
tail -n 4 gpurun_out/dp8_*.err | tail -40
