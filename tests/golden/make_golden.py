#!/usr/bin/env python3
"""Extract the reference's golden checksums for the hot path into tests/golden/.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference's goldens are the pinned known answers for the oracle:
  tests/checksum/benchmarks_json/linear_wake.normalized.1Rank.json
  tests/checksum/benchmarks_json/blowout_wake_explicit.2Rank.json
  tests/checksum/benchmarks_json/beam_evolution.1Rank.json
  tests/checksum/benchmarks_json/laser_blowout_wake_explicit.SI.1Rank.json  (SI units + laser, step 0)
Their decks + overrides (tests/linear_wake.normalized.1Rank.sh:32-35,
tests/blowout_wake_explicit.2Rank.sh:32-35) are restated in examples/ of this repo.
"""
import json, os
REF = '/root/reference/tests/checksum/benchmarks_json'
OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {
    'linear_wake.normalized.1Rank': dict(
        deck='examples/linear_wake_normalized.in',
        overrides={'diagnostic.field_data': 'all rho'}, rtol_cpu=1e-12, rtol_cuda=1e-7),
    'blowout_wake_explicit.2Rank': dict(
        deck='examples/blowout_wake_normalized.in',
        overrides={'max_step': 1}, rtol_cpu=1e-9, rtol_cuda=1e-9,
        skip=['Sy', 'Sx', 'chi']),
    # tests/beam_evolution.1Rank.sh:33-44 (hipace.tile_size / output_period / file_prefix dropped)
    'beam_evolution.1Rank': dict(
        deck='examples/beam_in_vacuum_normalized.in',
        overrides={'amr.n_cell': '32 32 10', 'max_step': 20, 'geometry.prob_lo': '-2. -2. -2.',
                   'geometry.prob_hi': '2. 2. 2.', 'hipace.dt': 3., 'beam.density': 1.e-8,
                   'beam.radius': 1., 'beam.ppc': '4 4 1',
                   'beams.external_E(x,y,z,t)': '.5*x .5*y 0.'},
        rtol_cpu=1e-12, rtol_cuda=2e-6),
    # tests/laser_blowout_wake_explicit.SI.1Rank.sh:23-36: SI units, no beam, gaussian laser; pins
    # the SI code path, the laser initialisation, |a|^2 on the field grid and the ponderomotive terms
    # of deposit / explicit deposition / push at time step 0 (the envelope advance is not exercised)
    'laser_blowout_wake_explicit.SI.1Rank': dict(
        deck='examples/blowout_wake_SI.in',
        overrides={'max_step': 0, 'beams.names': 'no_beam',
                   'geometry.prob_lo': '-20.*kp_inv -20.*kp_inv -7.5*kp_inv',
                   'geometry.prob_hi': '20.*kp_inv 20.*kp_inv 6.*kp_inv',
                   'lasers.names': 'laser', 'lasers.lambda0': '.8e-6', 'laser.a0': 4.5,
                   'laser.position_mean': '0. 0. 0', 'laser.w0': '4.*kp_inv', 'laser.L0': '2.*kp_inv',
                   'amr.n_cell': '128 128 100'},
        rtol_cpu=1e-9, rtol_cuda=1e-9, skip=['Sy', 'Sx', 'chi']),
    # tests/laser_blowout_wake_explicit.1Rank.sh:23-36: the same in normalised units
    'laser_blowout_wake_explicit.1Rank': dict(
        deck='examples/blowout_wake_normalized.in',
        overrides={'max_step': 0, 'beams.names': 'no_beam', 'geometry.prob_lo': '-20. -20. -7.5',
                   'geometry.prob_hi': '20. 20. 6', 'lasers.names': 'laser', 'lasers.lambda0': '.8e-6',
                   'laser.a0': 4.5, 'laser.position_mean': '0. 0. 0', 'laser.w0': 4, 'laser.L0': 2,
                   'amr.n_cell': '128 128 100'},
        rtol_cpu=1e-9, rtol_cuda=1e-9, skip=['Sy', 'Sx', 'chi']),
    # tests/laser_evolution.SI.2Rank.sh:29-44: the envelope ADVANCE over 31 time steps in vacuum (the
    # checksummed output is the one of the fft solver run), xz diagnostic
    'laser_evolution.SI.2Rank': dict(
        deck='examples/laser_vacuum_SI.in', overrides={'lasers.solver_type': 'fft'},
        rtol_cpu=1e-12, rtol_cuda=1e-7),
    # tests/linear_wake.SI.1Rank.sh:30-34: SI units with a beam and the rho diagnostic
    'linear_wake.SI.1Rank': dict(
        deck='examples/linear_wake_SI.in', overrides={'diagnostic.field_data': 'all rho'},
        rtol_cpu=1e-11, rtol_cuda=1e-7),
}
for name, meta in CASES.items():
    gold = json.load(open(os.path.join(REF, name + '.json')))
    json.dump(dict(meta, source=f'tests/checksum/benchmarks_json/{name}.json', checksums=gold),
              open(os.path.join(OUT, name + '.json'), 'w'), indent=1, sort_keys=True)
    print('wrote', name)
