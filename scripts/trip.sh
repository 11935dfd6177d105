mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_dp_gpu.py -q -m gpu -s --no-header -p no:cacheprovider -k two_rank > gpurun_out/test_dp_gpu.log 2>&1; echo "dp exit=$?"
grep -h "^param\|replica\|passed\|failed\|Error" gpurun_out/test_dp_gpu.log | head -30
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --mode train --steps 20 --warmup 5 > gpurun_out/bench_train_2gpu.json 2> gpurun_out/bench_train_2gpu.err; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_train_2gpu.json | head -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_render_2gpu.json 2> gpurun_out/bench_render_2gpu.err; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_render_2gpu.json | head -3
