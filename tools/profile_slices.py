#!/usr/bin/env python3
"""Run the slice loop down to a given slice, then bracket a few slices with cudaProfilerStart/Stop
so that `ncu --profile-from-start off` captures representative (blow-out) slices only.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_slices.py --skip 400 --slices 2
Without ncu it prints the per-stage device time of the bracketed slices.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument('--nxy', type=int, default=1024)
ap.add_argument('--nz', type=int, default=1024)
ap.add_argument('--ppc', type=int, default=2)
ap.add_argument('--skip', type=int, default=400)
ap.add_argument('--slices', type=int, default=2)
args = ap.parse_args()

import torch
import hipace_b200 as hp
from bench import deck_and_overrides

deck, ov = deck_and_overrides(args.nxy, args.nz, args.ppc)
sim = hp.Simulation(deck, ov)
sim.set_option('checksums', 0)
sim.begin_step(0)
isl = args.nz - 1
for _ in range(args.skip):
    sim.solve_one_slice(isl)
    isl -= 1
torch.cuda.synchronize()
torch.cuda.profiler.start()
t0 = torch.cuda.Event(enable_timing=True)
for _ in range(args.slices):
    sim.solve_one_slice(isl)
    isl -= 1
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled slices', isl + args.slices, '..', isl + 1)
sim.close()
