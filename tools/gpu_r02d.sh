#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
for combo in "push_variant=0 expl_variant=4"; do
  timeout 300 python tools/debug_cta.py 256 512 "$combo" 2>&1 | tail -2
  timeout 300 python tools/debug_cta.py 1024 400 "$combo" 2>&1 | tail -2
done | tee gpurun_out/${tag}_debug.txt
timeout 1200 python -m pytest tests -m gpu -q -rfEs --tb=short --timeout 300 2>&1 | tail -60 > gpurun_out/${tag}_pytest.txt
tail -5 gpurun_out/${tag}_pytest.txt
timeout 900 bash tools/tune.sh "-" "push_variant=2 expl_variant=0" "push_variant=5" "push_variant=4" "expl_variant=8" "expl_variant=7" "push_variant=5 expl_variant=8" 2>&1 | tee gpurun_out/${tag}_tune.txt
timeout 300 python bench.py 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl cufft_ref --no-cpu-baseline --no-e2e 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_cufft_ref.json
timeout 400 python bench.py --impl naive --no-cpu-baseline --no-e2e --steps 2 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_naive.json
ls -la gpurun_out | tail -8
