#!/bin/bash
# Evidence trip (1 GPU): whole GPU test-suite, bench lines (default + the regimes that leave L2), ncu launch lists and full
# captures of the kernels that ship at HEAD.  usage: trip_evidence.sh <tag>
R=${1:-r2f}
mkdir -p gpurun_out
bash scripts/gpu_tests.sh
timeout 900 python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err; tail -c 600 gpurun_out/${R}_bench_default.json
for seg in 100 100,100,50; do
  tag=$(echo $seg | tr ',' '-')
  HRF_BENCH_SEGMENTS=$seg timeout 300 python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline --no-companions > gpurun_out/${R}_bench_train_seg$tag.json 2> gpurun_out/${R}_bench_train_seg$tag.err
  HRF_BENCH_SEGMENTS=$seg timeout 300 python bench.py --mode render --steps 20 --warmup 5 --no-cpu-baseline --no-companions > gpurun_out/${R}_bench_render_seg$tag.json 2> gpurun_out/${R}_bench_render_seg$tag.err
  python -c "
import json
for m in ('train','render'):
    try:
        l=json.loads(open('gpurun_out/${R}_bench_%s_seg$tag.json' % m).readline()); print('seg $seg', m, round(l['value']), 'rays/s', round(l['ms_per_step'],3), 'ms', 'kernel_ms', round(l['roofline']['kernel_ms'],3), l.get('phases_ms'))
    except Exception as e: print('seg $seg', m, 'failed', e)
"
done
python scripts/kernel_times.py --segments 100 100 50 --reps 5 > gpurun_out/${R}_kernel_times_seg100-100-50.txt 2>&1
python scripts/kernel_times.py --segments 50 > gpurun_out/${R}_kernel_times_seg50.txt 2>&1
HRF_SCATTER_CTAS=6 HRF_BWD_UNROLL=1 python scripts/kernel_times.py --segments 50 2>&1 | grep -i "scatter\|backward MLP" | sed "s/^/v3-6ctas, bwd unroll 1: /" > gpurun_out/${R}_kernel_times_variants.txt
python scripts/seed_spread.py > gpurun_out/${R}_seed_spread.txt 2>&1
python scripts/image_phases.py > gpurun_out/${R}_image_phases.txt 2>&1; tail -1 gpurun_out/${R}_image_phases.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_train.csv \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_l_train.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${R}_launches_render.csv \
    python bench.py --mode render --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_l_render.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"field_forward|field_backward|grid_scatter|adam_multi" -s 15 -c 5 -o gpurun_out/prof_${R}_train -f \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_f_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"field_forward" -s 3 -c 1 -o gpurun_out/prof_${R}_render -f \
    python bench.py --mode render --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_f_render.log 2>&1
HRF_BENCH_SEGMENTS=100,100,50 timeout 900 ncu --set full --clock-control none -k regex:"field_forward" -s 3 -c 1 -o gpurun_out/prof_${R}_render_seg250 -f \
    python bench.py --mode render --steps 2 --warmup 3 --no-cpu-baseline --no-companions > gpurun_out/ncu_f_render250.log 2>&1
ls -la gpurun_out/prof_${R}_*.ncu-rep
