#!/bin/bash
tag=${1:-r02P}
mkdir -p gpurun_out
for combo in "mg_persist=0" "mg_persist=1" "mg_persist=2"; do
  timeout 120 python tools/debug_cta.py 256 256 "$combo" 2>&1 | tail -1
  timeout 120 python tools/debug_cta.py 1024 64 "$combo" 2>&1 | tail -1
  timeout 120 python tools/debug_cta.py 1023 48 "$combo" 2>&1 | tail -1
done | tee gpurun_out/${tag}_debug.txt
export TUNE_ARGS="--steps 3"
timeout 600 bash tools/tune.sh "mg_persist=0" "mg_persist=1" "mg_persist=2" "mg_persist=0" "mg_persist=2" 2>&1 | tee gpurun_out/${tag}_tune.txt
export TUNE_ARGS="--steps 3 --workload configs1"
timeout 300 bash tools/tune.sh "mg_persist=0" "mg_persist=1" "mg_persist=2" 2>&1 | tee -a gpurun_out/${tag}_tune.txt
