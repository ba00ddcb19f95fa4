#!/bin/bash
tag=${1:-r02q}
mkdir -p gpurun_out
for combo in "push_variant=9 expl_variant=11" "push_variant=6 expl_variant=12"; do
  timeout 300 python tools/debug_cta.py 256 512 "$combo" 2>&1 | tail -1
done | tee gpurun_out/${tag}_debug.txt
timeout 300 python tools/debug_cta.py 256 512 "push_variant=6 expl_variant=0" 2>&1 | tail -1 | tee -a gpurun_out/${tag}_debug.txt
timeout 900 bash tools/tune.sh "-" "push_variant=9" "expl_variant=11" "expl_variant=12" "push_variant=9 expl_variant=12" 2>&1 | tee gpurun_out/${tag}_tune.txt
