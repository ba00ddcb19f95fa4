#!/bin/bash
tag=${1:-r02c}
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=1
for combo in "push_variant=0 expl_variant=0" "push_variant=2 expl_variant=4" "push_variant=4 expl_variant=0" "push_variant=2 expl_variant=7"; do
  timeout 200 python tools/debug_cta.py 256 512 "$combo" 2>&1 | tail -2
done | tee gpurun_out/${tag}_debug256.txt
for combo in "push_variant=0 expl_variant=0" "push_variant=2 expl_variant=4"; do
  timeout 200 python tools/debug_cta.py 1024 300 "$combo" 2>&1 | tail -2
done | tee gpurun_out/${tag}_debug1024.txt
