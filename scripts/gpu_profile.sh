#!/bin/bash
# ncu evidence for profiles/: launch lists of the bench commands + full captures of the dominant kernels.
mkdir -p gpurun_out
R=${1:-r1}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_${R}_render.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_render.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_${R}_train.csv \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_forward -s 3 -c 1 -o gpurun_out/prof_${R}_fwd -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"field_backward|grid_scatter" -s 6 -c 2 -o gpurun_out/prof_${R}_bwd -f \
    python bench.py --mode train --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_* 2>&1
