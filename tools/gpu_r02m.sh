#!/bin/bash
tag=${1:-r02m}
mkdir -p gpurun_out
export HPB_BENCH_WATCHDOG=600
timeout 400 python bench.py --no-cpu-baseline 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json | cut -c1-250
timeout 400 python bench.py --workload n1023 --no-cpu-baseline --no-e2e 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_n1023.json | cut -c1-250
timeout 600 python bench.py --workload configs3 --steps 1 --warmup 3 --no-cpu-baseline 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_configs3.json | cut -c1-250
timeout 900 python bench.py --workload configs4 --steps 1 --warmup 3 --no-cpu-baseline 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_configs4.json | cut -c1-250
tail -5 gpurun_out/${tag}_bench.err
