#!/usr/bin/env bash
# One `ncu --set full` capture of the fused renderer's kernel inside the default bench command + the launch list of the same
# command (B200_PROFILING.md recipe).  usage: ncu_ws3.sh <tag>   ->  gpurun_out/<tag>.ncu-rep, gpurun_out/<tag>_launches.csv
tag=$1
ncu --set full --clock-control none --import-source on -k regex:k_render_ws3 -s 6 -c 1 -f -o gpurun_out/$tag \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-variants --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
