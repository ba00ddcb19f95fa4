#!/bin/bash
# prime FFT stage on the fp64 tensor cores (DMMA.8x8x4) vs the scalar inner product (fft_variant=2)
tag=${1:-r02z}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_periodic.py tests/test_gpu_zzz_late_features.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/${tag}_pytest.txt
export TUNE_ARGS="--steps 3"
timeout 900 bash tools/tune.sh "-" "fft_variant=2" "-" "fft_variant=2" 2>&1 | tee gpurun_out/${tag}_tune.txt
