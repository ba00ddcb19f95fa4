#!/bin/bash
tag=${1:-r02o}
mkdir -p gpurun_out
export HPB_BENCH_WATCHDOG=800
timeout 900 python bench.py --workload configs4 --steps 1 --warmup 3 --no-cpu-baseline 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_configs4.json | cut -c1-250
tail -3 gpurun_out/${tag}_bench.err
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
