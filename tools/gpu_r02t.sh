#!/bin/bash
tag=${1:-r02t}
mkdir -p gpurun_out
for combo in "expl_variant=13" "expl_variant=15"; do
  timeout 300 python tools/debug_cta.py 256 512 "$combo" 2>&1 | tail -1
done | tee gpurun_out/${tag}_debug.txt
timeout 900 bash tools/tune.sh "-" "expl_variant=13" "expl_variant=14" "expl_variant=15" 2>&1 | tee gpurun_out/${tag}_tune.txt
