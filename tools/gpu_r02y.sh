#!/bin/bash
# multigrid launch-chain experiments: level-0 buffer rotation (no final copy) and the coefficient hierarchy
# on a third stream beside the Poisson solve / explicit deposition
tag=${1:-r02y}
mkdir -p gpurun_out
for combo in "mg_rotate=0 mg_early=0" "mg_rotate=1 mg_early=0" "mg_rotate=0 mg_early=1" "mg_rotate=1 mg_early=1"; do
  timeout 300 python tools/debug_cta.py 256 512 "$combo" 2>&1 | tail -1
done | tee gpurun_out/${tag}_debug.txt
timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mg_solve1 or slice_by_slice" 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest.txt
export TUNE_ARGS="--steps 4"
timeout 900 bash tools/tune.sh "mg_rotate=0 mg_early=0" "mg_rotate=1 mg_early=0" "mg_rotate=0 mg_early=1" "mg_rotate=1 mg_early=1" "mg_rotate=0 mg_early=0" "mg_rotate=1 mg_early=1" 2>&1 | tee gpurun_out/${tag}_tune.txt
