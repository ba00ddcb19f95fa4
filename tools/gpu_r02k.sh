#!/bin/bash
tag=${1:-r02k}
mkdir -p gpurun_out
export HPB_BENCH_WATCHDOG=400
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/${tag}_bench2.err | tee gpurun_out/${tag}_bench2_1ch.json | cut -c1-400
NCCL_MAX_P2P_NCHANNELS=32 NCCL_MIN_P2P_NCHANNELS=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-verify 2>> gpurun_out/${tag}_bench2.err | tee gpurun_out/${tag}_bench2_default.json | cut -c1-400
