#!/bin/bash
# ncu --set full capture of ONE window-attention launch (stage / window / impl from the arguments); report lands in gpurun_out/
# usage: tools/ncu_attn.sh <tag> <impl> <stage> <w7|w12> <shift 0|1>
mkdir -p gpurun_out
LAVT_ATTN_IMPL=$2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn -s 4 -c 1 -f \
  -o gpurun_out/prof_attn_$1 python tools/run_attn_once.py $3 $4 $5 > gpurun_out/ncu_attn_$1.log 2>&1
tail -2 gpurun_out/ncu_attn_$1.log
