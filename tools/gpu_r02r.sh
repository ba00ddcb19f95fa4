#!/bin/bash
tag=${1:-r02r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfEs --tb=short --timeout 400 -k "poisson or golden or full_size or laser" 2>&1 | tail -8 | tee gpurun_out/${tag}_pytest.txt
timeout 900 bash tools/tune.sh "-" "mg_wide=1" "mg_fuse=1" "fft_variant=1" 2>&1 | tee gpurun_out/${tag}_tune.txt
