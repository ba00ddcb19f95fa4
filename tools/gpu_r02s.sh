#!/bin/bash
# 4-GPU visit: the default line and BASELINE configs[3] (laser, multigrid envelope solver) on 4 ranks
tag=${1:-r02s}
mkdir -p gpurun_out
export HPB_BENCH_WATCHDOG=500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/${tag}_bench4.err | tee gpurun_out/${tag}_bench4.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --workload configs3 --steps 2 --warmup 3 --no-cpu-baseline 2>> gpurun_out/${tag}_bench4.err | tee gpurun_out/${tag}_bench4_configs3.json | cut -c1-300
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${tag}_bench4.err | tail -5
