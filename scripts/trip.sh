mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
CHOICE=0
if timeout 120 python -m pytest tests/test_scatter_gpu.py -q -m gpu -s --no-header -p no:cacheprovider -k tapstage > gpurun_out/test_scatter_tapstage.log 2>&1; then
  for c in 0 1; do HRF_SCATTER_TAPSTAGE=$c timeout 100 python scripts/train_phases.py 2>&1 | grep flush | tail -1 | sed "s/^/tapstage $c: /"; done | tee gpurun_out/scatter_tapstage.txt
  CHOICE=$(python - <<'PY'
import re
t=open("gpurun_out/scatter_tapstage.txt").read()
v=[float(x) for x in re.findall(r"'backward': ([0-9.]+)", t)]
print(1 if len(v)==2 and v[1] < v[0] else 0)
PY
)
else
  echo "tapstage variant FAILED its test"; tail -5 gpurun_out/test_scatter_tapstage.log
fi
echo "HRF_SCATTER_TAPSTAGE=$CHOICE" | tee gpurun_out/tapstage_choice.txt
export HRF_SCATTER_TAPSTAGE=$CHOICE
timeout 700 bash scripts/gpu_tests.sh > gpurun_out/gpu_tests.out 2>&1; cat gpurun_out/summary.txt
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_render.json 2> gpurun_out/bench_render.err; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/bench_render.json | head -5
timeout 120 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/bench_train.json | head -5
timeout 120 python bench.py --mode image --steps 3 --warmup 3 > gpurun_out/bench_image.json 2> gpurun_out/bench_image.err; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_image.json | head -3
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
