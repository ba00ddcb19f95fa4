#!/bin/bash
# round 2, first GPU visit: the strict GPU suite (no xfail marks any more), the default bench line and the
# two extra sizes, the option order=9 A/B the round-1 verdict asked for.
tag=${1:-r02a}
mkdir -p gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket|Core|Thread"; nvidia-smi -L; } > gpurun_out/${tag}_host.txt
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
timeout 1200 python -m pytest tests -m gpu -q -rfEs --tb=short 2>&1 | tail -150 > gpurun_out/${tag}_pytest.txt
tail -5 gpurun_out/${tag}_pytest.txt
timeout 600 python bench.py 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_reference.json
timeout 300 python bench.py --workload n1023 --no-cpu-baseline 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_n1023.json
timeout 300 python bench.py --workload configs1 --steps 6 --no-cpu-baseline 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_configs1.json
timeout 400 bash tools/tune.sh "-" "order=9" "order=3" "order=0" 2>&1 | tee gpurun_out/${tag}_tune.txt
ls -la gpurun_out | tail -12
