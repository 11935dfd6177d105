#!/bin/bash
# Round 2, last trip: HEAD sanity (defaults) -- scatter / training / baseline-size parity, smoke(), default bench line.
mkdir -p gpurun_out
: > gpurun_out/summary13.txt
for t in tests/test_scatter_gpu.py tests/test_training_gpu.py tests/test_baseline_sizes_gpu.py; do
  n=$(basename $t .py)
  timeout 900 python -m pytest $t -q -m gpu -s -x --no-header -p no:cacheprovider > gpurun_out/${n}_13.log 2>&1
  echo "${n} exit=$? $(tail -1 gpurun_out/${n}_13.log)" | tee -a gpurun_out/summary13.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/summary13.txt
timeout 900 python bench.py > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err; tail -c 300 gpurun_out/r2n_bench_default.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err; tail -c 400 gpurun_out/r2n_bench_reference.json
