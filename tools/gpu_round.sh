#!/bin/bash
# One GPU visit: smoke, parity tests, the default bench line (both arms), the launch list of two
# blow-out slices and a full ncu capture of the particle kernels.  Output: gpurun_out/<tag>_* (small).
tag=${1:-r01}
mkdir -p gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket|Core|Thread"; nvidia-smi -L; } > gpurun_out/${tag}_host.txt
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q -rxX 2>&1 | tail -25 | tee gpurun_out/${tag}_pytest.txt
timeout 600 python bench.py 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_reference.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_slices.py --skip 400 --slices 2 2>&1 | tail -1
timeout 400 ncu --profile-from-start off --set full --clock-control none -f -o /tmp/${tag}_p \
    -k regex:'k_advance_plasma|k_explicit_deposition' -c 2 \
    python tools/profile_slices.py --skip 400 --slices 1 2>&1 | tail -1
ncu -i /tmp/${tag}_p.ncu-rep --page raw --csv > gpurun_out/${tag}_particles_raw.csv
ls -la gpurun_out
