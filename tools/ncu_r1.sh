#!/bin/bash
# ncu passes for the round-1 profile (run under gpurun; outputs land in gpurun_out/)
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
# (1) launch list of the bench command: eager warm-up forward + 3 warm-up steps are skipped (4 x 244 launches + copies)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 500 --csv \
    --log-file gpurun_out/launches_r1.csv $BENCH > gpurun_out/ncu_bench1.log 2>&1
# (2) DRAM traffic + tensor-pipe activity of every GEMM/conv launch of one step
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:gemm_bf16_tc -s 520 -c 130 --csv --log-file gpurun_out/gemm_traffic_r1.csv $BENCH > gpurun_out/ncu_bench2.log 2>&1
# (3) full-set captures: a few GEMM/conv launches and the attention kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 640 -c 6 \
    -o gpurun_out/prof_gemm_r1 -f $BENCH > gpurun_out/ncu_bench3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn -s 96 -c 3 \
    -o gpurun_out/prof_attn_r1 -f $BENCH > gpurun_out/ncu_bench4.log 2>&1
tail -3 gpurun_out/ncu_bench*.log
ls -la gpurun_out
