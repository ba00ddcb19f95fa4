#!/bin/bash
tag=${1:-r02I}
mkdir -p gpurun_out
for combo in "push_variant=10"; do
  timeout 300 python tools/debug_cta.py 256 512 "$combo" 2>&1 | tail -1
done | tee gpurun_out/${tag}_debug.txt
export TUNE_ARGS="--steps 3"
timeout 900 bash tools/tune.sh "-" "push_variant=10" "-" "push_variant=10" 2>&1 | tee gpurun_out/${tag}_tune.txt
