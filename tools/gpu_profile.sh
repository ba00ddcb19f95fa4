#!/bin/bash
# ncu captures of one blow-out slice (slice 623 of the 1024^3 deck); keeps gpurun_out/ small:
#  <tag>_launches.csv   per-launch durations of two slices
#  <tag>_raw.csv        --set full raw page of every kernel of one slice (no source)
#  <tag>_hot.ncu-rep    --set full + source of the particle kernels and the level-0 smoother
tag=${1:-r01}
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_slices.py --skip 400 --slices 2 2>&1 | tail -1
ncu --profile-from-start off --set full --clock-control none -f -o /tmp/${tag}_slice \
    python tools/profile_slices.py --skip 400 --slices 1 2>&1 | tail -1
ncu -i /tmp/${tag}_slice.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:'k_advance_plasma|k_explicit_deposition|k_deposit_current|k_thomas_local' -c 4 -o gpurun_out/${tag}_hot \
    python tools/profile_slices.py --skip 400 --slices 1 2>&1 | tail -1
ls -la gpurun_out /tmp/${tag}_slice.ncu-rep
