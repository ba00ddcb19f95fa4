#!/bin/bash
# The evidence visit of a round: smoke, the whole GPU suite, the default bench line (both arms), the same-box
# GPU bars, the other BASELINE workloads, the launch list of two blow-out slices and full ncu captures of the
# hot kernels of the CURRENT defaults.  Output: gpurun_out/<tag>_* (small; raw csv pages instead of .ncu-rep).
tag=${1:-r02F}
mkdir -p gpurun_out
export HPB_BENCH_WATCHDOG=600
{ nproc; lscpu | grep -E "Model name|Socket|Core|Thread"; nvidia-smi -L; } > gpurun_out/${tag}_host.txt
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
timeout 900 python -m pytest tests -m gpu -q -rxXs 2>&1 | tail -12 | tee gpurun_out/${tag}_pytest.txt
timeout 600 python bench.py 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_reference.json | cut -c1-200
timeout 300 python bench.py --impl cufft_ref --no-cpu-baseline --no-e2e 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_cufft_ref.json | cut -c1-200
timeout 300 python bench.py --impl naive --no-cpu-baseline --no-e2e 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_naive.json | cut -c1-200
timeout 300 python bench.py --workload configs1 --no-cpu-baseline --no-e2e 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_configs1.json | cut -c1-200
timeout 300 python bench.py --workload n1023 --no-cpu-baseline --no-e2e 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_n1023.json | cut -c1-200
timeout 300 bash tools/tune.sh "-" 2>&1 | tee gpurun_out/${tag}_tune.txt
timeout 400 python bench.py --workload configs3 --steps 1 --warmup 3 --no-cpu-baseline 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_configs3.json | cut -c1-200
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_slices.py --skip 400 --slices 2 2>&1 | tail -1
timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o /tmp/${tag}_p \
    -k regex:'k_advance_plasma|k_explicit_deposition' -c 2 \
    python tools/profile_slices.py --skip 400 --slices 1 2>&1 | tail -1
ncu -i /tmp/${tag}_p.ncu-rep --page raw --csv > gpurun_out/${tag}_particles_raw.csv
timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o /tmp/${tag}_s \
    -k regex:'k_dst_rows|k_smooth|k_thomas' -c 7 \
    python tools/profile_slices.py --skip 400 --slices 1 2>&1 | tail -1
ncu -i /tmp/${tag}_s.ncu-rep --page raw --csv > gpurun_out/${tag}_solvers_raw.csv
cp /tmp/${tag}_p.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out | grep ${tag}
