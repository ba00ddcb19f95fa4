#!/bin/bash
tag=${1:-r02f}
mkdir -p gpurun_out
for combo in "push_variant=2 expl_variant=0" "push_variant=6 expl_variant=0" "push_variant=7 expl_variant=0" "push_variant=8 expl_variant=0"; do
  timeout 300 python tools/debug_cta.py 256 512 "$combo" 2>&1 | tail -1
done | tee gpurun_out/${tag}_debug.txt
timeout 300 python tools/debug_cta.py 1023 200 "push_variant=2 expl_variant=0" "push_variant=6 expl_variant=0" 2>&1 | tail -2 | tee -a gpurun_out/${tag}_debug.txt
timeout 900 bash tools/tune.sh "push_variant=2 expl_variant=0" "push_variant=6 expl_variant=0" "push_variant=7 expl_variant=0" "push_variant=8 expl_variant=0" "push_variant=5 expl_variant=0" 2>&1 | tee gpurun_out/${tag}_tune.txt
timeout 200 python -m pytest tests/test_gpu_zzz_late_features.py -m gpu -q -x --timeout 300 -k "laser" 2>&1 | tail -5 | tee gpurun_out/${tag}_laser_pytest.txt
