mkdir -p gpurun_out
for f in tests/test_dp_gpu.py tests/test_occupancy_tools_gpu.py; do
  timeout 500 python -m pytest $f -q -m gpu -s --no-header -p no:cacheprovider > gpurun_out/$(basename $f .py).log 2>&1; echo "$f exit=$?"; tail -2 gpurun_out/$(basename $f .py).log
done
grep -h "carve vs\|^param\|replicas" gpurun_out/*.log
