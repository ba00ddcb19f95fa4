#!/bin/bash
# First GPU visit of a round (ROADMAP.md section 1): the whole GPU suite with the outcome of every
# late-feature test spelled out, then the default bench line.  Output: gpurun_out/<tag>_first_*.
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rxXfE --tb=short 2>&1 | tail -120 > gpurun_out/${tag}_first_pytest.txt
tail -30 gpurun_out/${tag}_first_pytest.txt
timeout 600 python bench.py --steps 2 --warmup 3 2> gpurun_out/${tag}_first_bench.err | tee gpurun_out/${tag}_first_bench.json
# the experiment of ROADMAP.md section 2.1
timeout 300 bash tools/tune.sh "HPB_ORDER=1" "HPB_ORDER=9" "HPB_ORDER=3" 2>&1 | tee gpurun_out/${tag}_first_tune.txt
