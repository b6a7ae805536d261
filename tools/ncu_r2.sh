#!/bin/bash
# ncu pass behind profiles/r2_dram_traffic.json (run under gpurun; outputs land in gpurun_out/): DRAM bytes, duration and tensor-pipe
# activity of every gemm_bf16_tc_kernel launch of ONE eager bench step (130 launches per step; the first 4 x 130 are warm-up / graph
# capture passes of `bench.py --steps 2 --warmup 3 --no-graph`), then tools/dram_traffic_json.py turns the CSV into the JSON bench.py reads.
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:gemm_bf16_tc -s 520 -c 130 --csv --log-file gpurun_out/r2_gemm_traffic.csv $BENCH > gpurun_out/ncu_r2_traffic.log 2>&1
tail -1 gpurun_out/ncu_r2_traffic.log | cut -c1-120
python tools/dram_traffic_json.py gpurun_out/r2_gemm_traffic.csv 130 > gpurun_out/r2_dram_traffic.json && cat gpurun_out/r2_dram_traffic.json
