#!/bin/bash
tag=${1:-r02v}
mkdir -p gpurun_out
export HPB_BENCH_WATCHDOG=600
timeout 600 python -m pytest tests -m gpu -q -rfEs --tb=short --timeout 400 -k "mg_solve2 or laser" 2>&1 | tail -8 | tee gpurun_out/${tag}_pytest.txt
timeout 600 python bench.py --workload configs3 --steps 1 --warmup 3 --no-cpu-baseline 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_configs3.json | cut -c1-250
tail -3 gpurun_out/${tag}_bench.err
