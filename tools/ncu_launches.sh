#!/bin/bash
# launch list (gpu__time_duration) of one eager bench step; usage: tools/ncu_launches.sh <tag> [extra bench args]
tag=$1; shift
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 500 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline "$@" > gpurun_out/ncu_launch_$tag.log 2>&1
tail -2 gpurun_out/ncu_launch_$tag.log
