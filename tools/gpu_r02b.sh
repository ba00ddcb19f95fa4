#!/bin/bash
# round 2, second GPU visit: the CTA-tile particle kernels (TMA patch prefetch + shared-memory combine).
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
timeout 1200 python -m pytest tests -m gpu -q -rfEs --tb=short -x --timeout 300 2>&1 | tail -60 > gpurun_out/${tag}_pytest.txt
tail -5 gpurun_out/${tag}_pytest.txt
timeout 900 bash tools/tune.sh "-" "push_variant=2 expl_variant=0" "push_variant=5" "push_variant=4" "expl_variant=8" "expl_variant=7" "push_variant=5 expl_variant=8" 2>&1 | tee gpurun_out/${tag}_tune.txt
timeout 300 python bench.py 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json
ls -la gpurun_out | tail -8
