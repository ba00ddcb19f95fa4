#!/bin/bash
# One GPU visit: parity tests, the default bench line, the launch list of two blow-out slices and a
# full ncu capture of one slice.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-r01}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/${tag}_host.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${tag}_pytest.txt
python bench.py 2> gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_slices.py --skip 400 --slices 2 2>&1 | tail -1
ncu --profile-from-start off --set full --clock-control none --import-source on -f \
    -o gpurun_out/${tag}_slice python tools/profile_slices.py --skip 400 --slices 1 2>&1 | tail -1
