#!/bin/bash
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfEs --tb=short --timeout 400 2>&1 | tail -70 > gpurun_out/${tag}_pytest.txt
tail -6 gpurun_out/${tag}_pytest.txt
timeout 900 bash tools/tune.sh "-" "expl_variant=9" "expl_variant=10" "expl_variant=4" 2>&1 | tee gpurun_out/${tag}_tune.txt
timeout 300 python bench.py 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json
