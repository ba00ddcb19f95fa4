#!/bin/bash
# 8-GPU visit: the default line and BASELINE configs[4] shape (2048^2 x 2048, ppc 9, two mobile species)
tag=${1:-r02u}
mkdir -p gpurun_out
export HPB_BENCH_WATCHDOG=700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/${tag}_bench8.err | tee gpurun_out/${tag}_bench8.json | cut -c1-300
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --workload configs4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-verify 2>> gpurun_out/${tag}_bench8.err | tee gpurun_out/${tag}_bench8_configs4.json | cut -c1-300
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/${tag}_bench8.err | tail -5
