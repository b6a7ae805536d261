#!/bin/bash
# ncu passes behind profiles/r1_*_final* (run under gpurun; outputs land in gpurun_out/)
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
# DRAM traffic + tensor-pipe activity of every GEMM/conv launch of one step (launches 0..129 of the first timed step)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:gemm_bf16_tc -s 520 -c 130 --csv --log-file gpurun_out/gemm_traffic_final.csv $BENCH > gpurun_out/ncu_final1.log 2>&1
# same for the attention launches of one step
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:window_attn -s 96 -c 24 --csv --log-file gpurun_out/attn_traffic_final.csv $BENCH > gpurun_out/ncu_final2.log 2>&1
tail -1 gpurun_out/ncu_final1.log | cut -c1-100; tail -1 gpurun_out/ncu_final2.log | cut -c1-100
