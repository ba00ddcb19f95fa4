#!/bin/bash
# full ncu capture (with source) of the kernels matching $2 in one blow-out slice: tools/gpu_profile_k.sh tag 'regex' count
tag=$1; rx=$2; cnt=${3:-8}
mkdir -p gpurun_out
ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -k regex:"$rx" -c $cnt -o gpurun_out/${tag} python tools/profile_slices.py --skip 400 --slices 1 2>&1 | tail -1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv
ls -la gpurun_out
