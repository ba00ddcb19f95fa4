#!/usr/bin/env python3
"""Find the slice at which a kernel variant combination fails: tools/debug_cta.py NXY NZ 'k=v k=v' ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hipace_b200 as hp
nxy, nz = int(sys.argv[1]), int(sys.argv[2])
deck = open('examples/blowout_wake_normalized.in').read()
ov = {'amr.n_cell': f'{nxy} {nxy} {nz}', 'plasma.ppc': '2 2'}
for combo in sys.argv[3:]:
    sim = hp.Simulation(deck, ov)
    sim.set_option('checksums', 1)
    for kv in combo.split():
        k, v = kv.split('=')
        sim.set_option(k, float(v))
    sim.begin_step(0)
    last = None
    try:
        for isl in range(nz - 1, -1, -1):
            sim.solve_one_slice(isl)
            last = isl
        cs = sim.checksums()
        print(combo, ': all', nz, 'slices ok;', ' '.join(f'{k}={cs[k]:.12e}' for k in ('Bx', 'ExmBy', 'Ez', 'jx', 'chi') if k in cs), flush=True)
    except Exception as e:
        print(combo, ': FAILED after slice', last, '->', str(e)[-160:], flush=True)
    try:
        sim.close()
    except Exception:
        pass
