#!/bin/bash
tag=${1:-r02J}
mkdir -p gpurun_out
for combo in "mg_lean=0" "mg_lean=1"; do
  timeout 300 python tools/debug_cta.py 256 512 "$combo" 2>&1 | tail -1
  timeout 300 python tools/debug_cta.py 1024 96 "$combo" 2>&1 | tail -1
done | tee gpurun_out/${tag}_debug.txt
export TUNE_ARGS="--steps 3"
timeout 900 bash tools/tune.sh "mg_lean=0" "mg_lean=1" "mg_lean=0" "mg_lean=1" 2>&1 | tee gpurun_out/${tag}_tune.txt
