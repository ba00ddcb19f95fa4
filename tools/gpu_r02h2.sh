#!/bin/bash
tag=${1:-r02H}
mkdir -p gpurun_out
for combo in "expl_variant=0" "expl_variant=16"; do
  timeout 300 python tools/debug_cta.py 256 512 "$combo" 2>&1 | tail -1
done | tee gpurun_out/${tag}_debug.txt
export TUNE_ARGS="--steps 3"
timeout 900 bash tools/tune.sh "-" "expl_variant=16" "-" "expl_variant=16" 2>&1 | tee gpurun_out/${tag}_tune.txt
