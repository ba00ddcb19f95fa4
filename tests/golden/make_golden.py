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
    # tests/gaussian_linear_wake.{normalized,SI}.1Rank.sh:32-43: gaussian fixed_ppc beam in the linear
    # wake decks; pin Sy, Sx and chi too (nothing skipped)
    'gaussian_linear_wake.normalized.1Rank': dict(
        deck='examples/linear_wake_normalized.in',
        overrides={'beam.profile': 'gaussian', 'beam.zmin': -5.9, 'beam.zmax': 5.9, 'beam.radius': 10,
                   'beam.position_mean': '0. 0. 0', 'beam.position_std': '2 2 1.41',
                   'geometry.prob_lo': '-10. -10. -6', 'geometry.prob_hi': '10. 10. 6',
                   'diagnostic.field_data': 'all rho'}, rtol_cpu=1e-12, rtol_cuda=1e-7),
    'gaussian_linear_wake.SI.1Rank': dict(
        deck='examples/linear_wake_SI.in',
        overrides={'beam.profile': 'gaussian', 'beam.zmin': '-59.e-6', 'beam.zmax': '59.e-6',
                   'beam.radius': '100.e-6', 'beam.position_mean': '0. 0. 0',
                   'beam.position_std': '20.e-6 20.e-6 14.1e-6',
                   'geometry.prob_lo': '-100.e-6 -100.e-6 -60.e-6',
                   'geometry.prob_hi': '100.e-6 100.e-6 60.e-6',
                   'diagnostic.field_data': 'all rho'}, rtol_cpu=1e-11, rtol_cuda=1e-7),
    # tests/beam_in_vacuum.{normalized,SI}.1Rank.sh:30-35: no plasma, ORDER-0 deposition, multigrid
    # Bx/By on a non-square 512 x 768 grid, tolerance 1e-5; the SI deck has two beam species
    'beam_in_vacuum.normalized.1Rank': dict(
        deck='examples/beam_in_vacuum_normalized.in',
        overrides={'hipace.depos_order_xy': 0, 'diagnostic.field_data': 'all rho',
                   'hipace.MG_tolerance_rel': 1e-5}, rtol_cpu=1e-12, rtol_cuda=3e-5),
    'beam_in_vacuum.SI.1Rank': dict(
        deck='examples/beam_in_vacuum_SI.in',
        overrides={'hipace.depos_order_xy': 0, 'diagnostic.field_data': 'all rho',
                   'hipace.MG_tolerance_rel': 1e-5}, rtol_cpu=1e-12, rtol_cuda=2.5e-5),
    # tests/grid_current.1Rank.sh:30-48: the analytic grid current added to jz_beam, order 0
    # (max_step = 1 with dt = 0: step 1 repeats step 0)
    'grid_current.1Rank': dict(
        deck='examples/beam_in_vacuum_normalized.in',
        overrides={'amr.n_cell': '32 32 32', 'max_step': 1, 'hipace.depos_order_xy': 0,
                   'geometry.prob_lo': '-8. -8. -6.', 'geometry.prob_hi': '8. 8. 6.',
                   'grid_current.use_grid_current': 1, 'grid_current.peak_current_density': 0.2,
                   'grid_current.position_mean': '0. 0. 0.',
                   'grid_current.position_std': '0.3 0.3 1.41', 'beam.profile': 'gaussian',
                   'beam.position_std': '0.3 0.3 1.41', 'beam.density': 0.2, 'beam.radius': 1.,
                   'beam.ppc': '1 1 1'}, rtol_cpu=1e-12, rtol_cuda=1e-5),
    # tests/beam_in_vacuum_open_boundary.normalized.1Rank.sh:30-44: the PREDICTOR-CORRECTOR Bx/By solver
    # with OPEN field boundaries (multipole expansion of the free-space potential, fields/OpenBoundary.H),
    # off-centre beam, order 0, absorbing particle boundary
    'beam_in_vacuum_open_boundary.normalized.1Rank': dict(
        deck='examples/beam_in_vacuum_normalized.in',
        overrides={'hipace.depos_order_xy': 0, 'hipace.bxby_solver': 'predictor-corrector',
                   'hipace.predcorr_B_mixing_factor': 0.95, 'hipace.predcorr_max_iterations': 5,
                   'boundary.field': 'Open', 'boundary.particle': 'Absorbing',
                   'geometry.prob_lo': '-4. -4. -2.', 'geometry.prob_hi': '4. 4. 2.',
                   'beam.position_mean': '2. -1. 0.', 'diagnostic.field_data': 'all rho'},
        rtol_cpu=1e-9, rtol_cuda=1e-9),
    # tests/adaptive_time_step.1Rank.sh:52-69 (the run whose output is checksummed: positive gradient):
    # hipace.dt = adaptive over 20 steps, the beam gaining energy in an external Ez = z / 2
    'adaptive_time_step.1Rank': dict(
        deck='examples/beam_in_vacuum_normalized.in',
        overrides={'amr.n_cell': '32 32 32', 'max_step': 20, 'geometry.prob_lo': '-2. -2. -2.',
                   'geometry.prob_hi': '2. 2. 2.', 'hipace.dt': 'adaptive', 'beam.density': 1,
                   'beam.radius': 1., 'beam.n_subcycles': 4, 'beam.ppc': '4 4 1',
                   'beams.external_E(x,y,z,t)': '0. 0. .5*z', 'plasmas.adaptive_density': 1,
                   'hipace.nt_per_betatron': 89.7597901025655},
        rtol_cpu=1e-12, rtol_cuda=1e-5),
}
for name, meta in CASES.items():
    gold = json.load(open(os.path.join(REF, name + '.json')))
    json.dump(dict(meta, source=f'tests/checksum/benchmarks_json/{name}.json', checksums=gold),
              open(os.path.join(OUT, name + '.json'), 'w'), indent=1, sort_keys=True)
    print('wrote', name)
