#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:'k_advance_plasma_cta|k_explicit_deposition_cta' -c 2 -o gpurun_out/${tag}_cta python tools/profile_slices.py --skip 400 --slices 1 2>&1 | tail -1
ncu -i gpurun_out/${tag}_cta.ncu-rep --page raw --csv > gpurun_out/${tag}_cta_raw.csv
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_slices.py --skip 400 --slices 2 2>&1 | tail -1
ls -la gpurun_out | tail -5
