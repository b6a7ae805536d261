#!/bin/bash
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/bench_train.py --gpus 2 --steps 5 --warmup 3 "$@" > gpurun_out/r2_t2_$tag.json 2> gpurun_out/r2_t2_$tag.err; echo "$tag $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_t2_$tag.json | head -1)"; }
run default
run noov --no-overlap-allreduce
run bf16 --bf16-allreduce
NCCL_MAX_CTAS=8 run cta8
NCCL_MAX_CTAS=4 run cta4
NCCL_MAX_CTAS=8 run cta8bf16 --bf16-allreduce
python tools/bench_train.py --steps 5 --warmup 3 > gpurun_out/r2_t2_single.json 2>/dev/null; echo "single $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_t2_single.json | head -1)"
