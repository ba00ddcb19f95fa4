#!/bin/bash
tag=${1:-r02l}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfEs --tb=short --timeout 400 2>&1 | tail -60 > gpurun_out/${tag}_pytest.txt
tail -6 gpurun_out/${tag}_pytest.txt
export TUNE_ARGS=""
timeout 900 bash tools/tune.sh "-" "bluestein_min_prime=5" "side_late=0" 2>&1 | tee gpurun_out/${tag}_tune.txt
timeout 300 python bench.py --workload configs1 --steps 6 --no-cpu-baseline 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_configs1.json | cut -c1-300
timeout 300 python bench.py --workload configs1 --steps 6 --no-cpu-baseline --no-e2e --opt bluestein_min_prime=1000 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_configs1_generic257.json | cut -c1-300
timeout 300 python bench.py --workload n1023 --no-cpu-baseline 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_n1023.json | cut -c1-300
