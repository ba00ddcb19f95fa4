"""In-situ beam diagnostics: the file our C++ writer produces is read by the
reference's own tools/read_insitu_diagnostics.py, holds what BeamParticleContainer::
InSituComputeDiags / InSituWriteToFile define, and the device-side per-particle terms (run on the
host) equal the oracle's."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle.hipace_oracle as O

REF_TOOLS = '/root/reference/tools'


def _synthetic_sums(ns=9, seed=0):
    rng = np.random.default_rng(seed)
    s = rng.uniform(0.5, 2.0, (23, ns))
    s[22] = rng.integers(1, 50, ns)
    s[:, 3] = 0.0               # an empty slice: sum(w) = 0 -> every average of that slice is 0
    return s


def test_writer_output_is_read_by_the_reference_tool(tmp_path):
    import hipace_b200 as hp
    path = tmp_path / 'reduced_beam.0000.txt'
    sums = _synthetic_sums()
    for step in (0, 1, 2):       # appended records, one header
        hp.insitu_write_beam(path, 0.25 * step, step, -1.0, 1.0, -6.0, 6.0, 0.01, True, sums * (1 + step))
    ours = hp.read_insitu(path)
    assert ours.shape == (3,) and list(ours['step']) == [0, 1, 2]
    for step in (0, 1, 2):
        dt, rec = O.insitu_beam_record(sums * (1 + step), 0.25 * step, step, -1.0, 1.0, -6.0, 6.0, 0.01, True)
        assert dt == ours.dtype
        assert ours[step].tobytes() == rec.tobytes()
    assert (ours['[x]'][:, 3] == 0).all() and (ours['Np'][:, 3] == 0).all()
    if not os.path.isdir(REF_TOOLS):
        pytest.skip('the reference tree is not here: checked with the restated reader only')
    sys.path.insert(0, REF_TOOLS)
    import read_insitu_diagnostics as R
    theirs = R.read_file(str(tmp_path / 'reduced_beam.*.txt'))
    assert theirs.dtype == ours.dtype and theirs.tobytes() == ours.tobytes()
    # and the derived quantities of that tool evaluate on it
    ex = R.emittance_x(theirs['average'])
    want = np.sqrt(np.abs((ours['average']['[x^2]'] - ours['average']['[x]'] ** 2)
                          * (ours['average']['[ux^2]'] - ours['average']['[ux]'] ** 2)
                          - (ours['average']['[x*ux]'] - ours['average']['[x]'] * ours['average']['[ux]']) ** 2))
    assert np.allclose(ex, want, rtol=1e-14)
    assert R.per_slice_charge(theirs).shape == (3, sums.shape[1]) if hasattr(R, 'per_slice_charge') else True


def test_device_terms_on_the_host_equal_the_oracle():
    from hipace_b200.build import build_host_check
    hc = C.CDLL(build_host_check())
    rng = np.random.default_rng(3)
    n = 5000
    for c in (1.0, 299792458.0):
        pc = O.PhysConst(c, 1., 1., 1., 1., 1836.)
        bs = {'x': rng.normal(0, 1, n), 'y': rng.normal(0, 1, n), 'z': rng.uniform(-1, 1, n),
              'w': rng.uniform(0.5, 2, n), 'ux': rng.normal(0, 3, n) * c, 'uy': rng.normal(0, 3, n) * c,
              'uz': rng.normal(100, 30, n) * c, 'valid': rng.uniform(0, 1, n) > 0.1}
        bs['uz'][::17] = 0.0               # uz = 0: the 1/uz terms drop out (:505)
        for radius in (np.inf, 1.5):
            want = O.beam_insitu_sums(bs, pc, radius)
            b7 = [np.ascontiguousarray(bs[k]) for k in ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz')]
            valid = np.ascontiguousarray(bs['valid'].astype(np.uint8))
            out = np.zeros(23)
            hc.hc_beam_insitu(C.c_long(n), (C.c_void_p * 7)(*[a.ctypes.data for a in b7]),
                              valid.ctypes.data_as(C.c_void_p), C.c_double(1.0 / c),
                              C.c_double(radius * radius), out.ctypes.data_as(C.c_void_p))
            assert out[22] == want[22] and 0 < want[22] < n
            assert np.abs(out - want).max() <= 1e-12 * np.abs(want).max()


def test_oracle_insitu_of_an_evolving_beam(repo_root):
    """beam_evolution deck, 3 steps: the per-slice and the slice-averaged moments are those of the
    beam before each push"""
    meta_ov = {'amr.n_cell': '32 32 10', 'max_step': 2, 'geometry.prob_lo': '-2. -2. -2.',
               'geometry.prob_hi': '2. 2. 2.', 'hipace.dt': 3., 'beam.density': 1.e-8, 'beam.radius': 1.,
               'beam.ppc': '4 4 1', 'beams.external_E(x,y,z,t)': '.5*x .5*y 0.', 'beams.insitu_period': 1}
    deck = open(os.path.join(repo_root, 'examples', 'beam_in_vacuum_normalized.in')).read()
    sim = O.Simulation(deck, meta_ov)
    seen = {}

    def hook(s, isl, stage):
        if stage == 'fields':           # before the push of this slice
            bs = s.beam_slice(s.beams[0], isl)
            n = bs['np']
            v = bs['valid'][:n]
            w = bs['w'][:n][v]
            seen[(s.step, isl)] = (w.sum(), (w * bs['x'][:n][v] ** 2).sum())
    sim.slice_hook = hook
    sim.evolve(step_end=2)
    recs = sim.insitu_records['beam']
    assert [int(r['step']) for r in recs] == [0, 1, 2]
    for r in recs:
        for isl in range(10):
            sw, swx2 = seen[(int(r['step']), isl)]
            assert r['sum(w)'][isl] == pytest.approx(sw, rel=1e-14)
            if sw > 0:
                assert r['[x^2]'][isl] == pytest.approx(swx2 / sw, rel=1e-13)
    # the focusing field makes the beam breathe: <x^2> changes from step to step
    assert recs[1]['average']['[x^2]'] != recs[0]['average']['[x^2]']


def _moment_floor(rec, nm):
    """absolute noise floor of the weighted moment `nm` of one in-situ record: first moments and
    cross terms of a symmetric beam vanish up to the round-off of sums whose terms have the size
    sqrt([a^2] [b^2]) -- the device reduces them in another order than the oracle, so they are held
    to 1e-10 of that scale, not of their own (round-off sized) value"""
    def rms(q):
        key = '[%s^2]' % q
        return float(np.sqrt(np.abs(rec[key]).max())) if key in rec.dtype.names else 0.
    body = nm.strip('[]')
    if '*' in body:
        a, b = body.split('*', 1)
        return 1e-10 * rms(a) * rms(b)
    if '/' in body:
        a, b = body.split('/', 1)
        m = float(np.abs(rec['[%s]' % b]).max()) if '[%s]' % b in rec.dtype.names else 0.
        return 1e-10 * rms(a) / m if m > 0 else 0.
    return 1e-10 * rms(body)


@pytest.mark.gpu
def test_cuda_insitu_files_match_oracle(repo_root, tmp_path):
    """the CUDA slice loop writes <prefix>/reduced_beam.0000.txt; same records as the oracle's"""
    import hipace_b200 as hp
    ov = {'amr.n_cell': '32 32 10', 'max_step': 2, 'geometry.prob_lo': '-2. -2. -2.',
          'geometry.prob_hi': '2. 2. 2.', 'hipace.dt': 3., 'beam.density': 1.e-8, 'beam.radius': 1.,
          'beam.ppc': '4 4 1', 'beams.external_E(x,y,z,t)': '.5*x .5*y 0.', 'beams.insitu_period': 1,
          'beams.insitu_file_prefix': str(tmp_path / 'insitu')}
    deck = open(os.path.join(repo_root, 'examples', 'beam_in_vacuum_normalized.in')).read()
    sim = hp.Simulation(deck, ov)
    sim.evolve(0, 2)
    sim.close()
    got = hp.read_insitu(tmp_path / 'insitu' / 'reduced_beam.0000.txt')
    ref = O.Simulation(deck, ov)
    ref.evolve(step_end=2)
    want = ref.insitu_records['beam']
    assert got.shape == (3,)
    for k, r in enumerate(want):
        assert got.dtype == r.dtype
        for nm in r.dtype.names:
            if nm in ('average', 'total'):
                for sub in r[nm].dtype.names:
                    assert got[k][nm][sub] == pytest.approx(r[nm][sub], rel=1e-10,
                                                            abs=_moment_floor(r[nm], sub) + 1e-300), (k, nm, sub)
            else:
                assert np.allclose(got[k][nm], r[nm], rtol=1e-10, atol=_moment_floor(r, nm) + 1e-300), (k, nm)


class _AdaptivePar(C.Structure):
    _fields_ = [(k, C.c_double) for k in ('nt_per_betatron', 'dt_max', 'threshold_uz', 'phase_tolerance')] + \
               [('phase_substeps', C.c_int), ('control_phase', C.c_int), ('c', C.c_double), ('ep0', C.c_double),
                ('numprocs', C.c_int), ('predict_step', C.c_int)]


@pytest.mark.parametrize('numprocs', [1, 2, 3])
def test_adaptive_time_step_host_logic_follows_the_oracle(numprocs, repo_root):
    """hpb_adaptive_dt_next (the C++ host arithmetic of hipace.dt = adaptive, csrc/adaptive_dt.hpp)
    fed with the per-step beam data of the oracle run that reproduces the reference's
    adaptive_time_step golden: the same sequence of time steps, bit for bit.  numprocs > 1: the
    bookkeeping of a pipeline of that many ranks (every rank's dt comes from ITS previous step and is
    used numprocs steps later, AdaptiveTimeStep.cpp:225-251)"""
    import hipace_b200 as hp
    import json
    meta = json.load(open(os.path.join(repo_root, 'tests', 'golden', 'adaptive_time_step.1Rank.json')))
    ov = dict(meta['overrides'], max_step=6)
    sim = O.Simulation(open(os.path.join(repo_root, meta['deck'])).read(), ov, numprocs=numprocs)
    log = []
    orig = sim._adaptive_from_min_uz

    def spy(t_now):
        ts = {k: dict(v) for k, v in sim._ts.items()}
        dt_in = sim.dt
        orig(t_now)
        log.append((ts, dt_in, sim.dt, sim._min_uz_mq))
    sim._adaptive_from_min_uz = spy
    dts = []
    sim.slice_hook = lambda s, isl, stage: dts.append(s.dt) if (stage == 'pushed' and isl == 0) else None
    sim.evolve(step_end=6)
    L = hp.lib()
    L.hpb_adaptive_dt_next.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_double] * 3 + [C.c_void_p] * 2
    par = _AdaptivePar(sim.nt_per_betatron, sim.dt_max, sim.adaptive_threshold_uz, sim.adaptive_phase_tolerance,
                       sim.adaptive_phase_substeps, 1, sim.pc.c, sim.pc.ep0, numprocs, 1)
    assert len(log) == 8 and len(dts) == 7            # the initial call + one per step
    for k, (ts, dt_in, dt_after_min_uz, mq_want) in enumerate(log):
        t = ts['beam']
        arr = np.array([t['min_uz'], t['sw'], t['swu'], t['swu2']])
        q, m = np.array([sim.beams[0].charge]), np.array([sim.beams[0].mass])
        dt_out, mq = C.c_double(0.), C.c_double(np.finfo(float).max)
        rc = L.hpb_adaptive_dt_next(C.byref(par), 1, arr.ctypes.data, q.ctypes.data, m.ctypes.data,
                                    abs(sim.adaptive_density * sim.pc.q_e), 0.0, dt_in, C.byref(dt_out), C.byref(mq))
        assert rc == 0
        assert mq.value == mq_want
        # uniform density: the phase-advance control leaves dt as CalculateFromMinUz set it
        assert dt_out.value == dt_after_min_uz
        # ... and that is the dt the owning rank's next step (numprocs steps later) ran with
        if k + numprocs - 1 < len(dts):
            assert dts[k + numprocs - 1] == dt_after_min_uz
    for k in range(min(numprocs, len(dts))):             # every rank starts from the broadcast initial estimate
        assert dts[k] == log[0][2]
    assert len(set(dts)) == len(dts) - (numprocs - 1)    # the step really adapts (after the ranks' common start)


def test_plasma_writer_and_terms(tmp_path, repo_root):
    """plasma in-situ diagnostics: writer vs the reference's reader and the oracle record; the
    device-side terms on the host vs the oracle; an oracle run for the content"""
    import hipace_b200 as hp
    from hipace_b200.build import build_host_check
    rng = np.random.default_rng(5)
    sums = rng.uniform(0.5, 2.0, (15, 8))
    sums[14] = rng.integers(1, 99, 8)
    sums[:, 5] = 0.0
    path = tmp_path / 'reduced_plasma.0000.txt'
    for step in (0, 1):
        hp.insitu_write_plasma(path, 0.5 * step, step, -1.0, 1.0, -6.0, 6.0, 0.02, True, sums * (1 + step))
    ours = hp.read_insitu(path)
    for step in (0, 1):
        dt, rec = O.insitu_plasma_record(sums * (1 + step), 0.5 * step, step, -1.0, 1.0, -6.0, 6.0, 0.02, True)
        assert dt == ours.dtype and ours[step].tobytes() == rec.tobytes()
    if os.path.isdir(REF_TOOLS):
        sys.path.insert(0, REF_TOOLS)
        import read_insitu_diagnostics as R
        theirs = R.read_file(str(tmp_path / 'reduced_plasma.*.txt'))
        assert theirs.tobytes() == ours.tobytes()
    # terms
    hc = C.CDLL(build_host_check())
    n = 4000
    pl = O.Plasma('p', -1.0, 1.0, (1, 1), None)
    pl.x, pl.y = rng.normal(0, 1, n), rng.normal(0, 1, n)
    pl.w, pl.psi = rng.uniform(0.5, 2, n), rng.uniform(0.3, 2, n)
    pl.ux, pl.uy = rng.normal(0, 0.7, n), rng.normal(0, 0.7, n)
    pl.valid = rng.uniform(0, 1, n) > 0.05
    for radius in (np.inf, 1.2):
        want = O.plasma_insitu_sums(pl, O.PhysConst.make(True), radius)
        r = [np.ascontiguousarray(a) for a in (pl.x, pl.y, pl.w, pl.ux, pl.uy, pl.psi)] + [np.zeros(n)] * 5
        valid = np.ascontiguousarray(pl.valid.astype(np.uint8))
        out = np.zeros(15)
        hc.hc_plasma_insitu(C.c_long(n), (C.c_void_p * 11)(*[a.ctypes.data for a in r]),
                            valid.ctypes.data_as(C.c_void_p), C.c_double(1.0), C.c_double(radius * radius),
                            out.ctypes.data_as(C.c_void_p))
        assert out[14] == want[14] and np.abs(out - want).max() <= 1e-12 * np.abs(want).max()
    # content: the unperturbed plasma ahead of the beam has <x> = 0, gamma = 1, no energy
    deck = open(os.path.join(repo_root, 'examples', 'blowout_wake_normalized.in')).read()
    sim = O.Simulation(deck, {'amr.n_cell': '32 32 100', 'plasmas.insitu_period': 1})
    sim.evolve(20)
    a = sim.plasma_insitu['plasma']     # raw per-slice sums [15, nz] (slices 99..80 filled)
    head = a[:, 99]
    assert head[14] == 32 * 32 and abs(head[11] / head[0] - 1.0) < 1e-14 and head[13] == 0.0
    assert a[13, 80] > 0.0              # the wake has given the plasma energy


@pytest.mark.gpu
def test_cuda_plasma_insitu_matches_oracle(repo_root, tmp_path):
    import hipace_b200 as hp
    deck = open(os.path.join(repo_root, 'examples', 'blowout_wake_normalized.in')).read()
    ov = {'amr.n_cell': '32 32 40', 'plasmas.insitu_period': 1,
          'plasmas.insitu_file_prefix': str(tmp_path / 'pl')}
    sim = hp.Simulation(deck, ov)
    sim.evolve(0, 0)
    sim.close()
    got = hp.read_insitu(tmp_path / 'pl' / 'reduced_plasma.0000.txt')
    ref = O.Simulation(deck, ov)
    ref.evolve(step_end=0)
    want = ref.plasma_insitu_records['plasma'][0]
    assert got.shape == (1,) and got.dtype == want.dtype
    for nm in want.dtype.names:
        if nm in ('average', 'total'):
            for sub in want[nm].dtype.names:
                assert got[0][nm][sub] == pytest.approx(want[nm][sub], rel=1e-10, abs=1e-12), (nm, sub)
        else:
            assert np.allclose(got[0][nm], want[nm], rtol=1e-10, atol=1e-12), nm


def test_field_insitu_writer_terms_and_oracle(tmp_path, repo_root):
    import hipace_b200 as hp
    from hipace_b200.build import build_host_check
    rng = np.random.default_rng(8)
    sums = rng.normal(0, 1, (10, 6))
    path = tmp_path / 'reduced_fields.0000.txt'
    for step in (0, 1):
        hp.insitu_write_fields(path, 0.5 * step, step, -6.0, 6.0, True, 0.03, sums * (1 + step))
    ours = hp.read_insitu(path)
    for step in (0, 1):
        dt, rec = O.insitu_field_record(sums * (1 + step), 0.5 * step, step, -6.0, 6.0, True, 0.03)
        assert dt == ours.dtype and ours[step].tobytes() == rec.tobytes()
    if os.path.isdir(REF_TOOLS):
        sys.path.insert(0, REF_TOOLS)
        import read_insitu_diagnostics as R
        assert R.read_file(str(tmp_path / 'reduced_fields.*.txt')).tobytes() == ours.tobytes()
    # device-side terms on the host against the oracle (SI magnitudes)
    hc = C.CDLL(build_host_check())
    geom = O.Geometry(20, 28, 8, (-1e-4, -1e-4, 0.0), (1e-4, 1.2e-4, 1e-5), 2, 2)
    pc = O.PhysConst.make(False)
    g = geom.g
    names = ('ExmBy', 'EypBx', 'Ez', 'Bx', 'By', 'Bz', 'jz_beam')
    F = {('This', nm): rng.normal(0, 1e9 if nm[0] == 'E' else (3.0 if nm[0] == 'B' else 1e12),
                                  (geom.ny + 2 * g, geom.nx + 2 * g)) for nm in names}
    want = O.field_insitu_sums(F, geom, pc)
    planes = np.ascontiguousarray(np.stack([F[('This', nm)] for nm in names]))
    sys.path.insert(0, os.path.dirname(__file__))
    from test_device_math_host import HcGrid
    hg = HcGrid(geom.nx + 2 * g, geom.ny + 2 * g, g, 0., 0., 1., 1.)
    out = np.zeros(10)
    hc.hc_field_insitu(planes.ctypes.data_as(C.c_void_p), C.byref(hg), (C.c_int * 7)(*range(7)), geom.nx, geom.ny,
                       C.c_double(pc.c), out.ctypes.data_as(C.c_void_p))
    assert np.abs(out - want).max() / np.abs(want).max() <= 1e-12
    assert (np.abs(out - want) <= 1e-11 * np.abs(want) + 1e-12 * np.abs(want).max()).all()
    # an oracle run: the beam loses energy to the wake, integral of Ez jz_beam < 0 for e- driver ... sign
    deck = open(os.path.join(repo_root, 'examples', 'blowout_wake_normalized.in')).read()
    sim = O.Simulation(deck, {'amr.n_cell': '32 32 100', 'fields.insitu_period': 1})
    sim.evolve(step_end=0)
    rec = sim.field_insitu_records[0]
    assert rec['integrated']['[Ez^2]'] > 0 and rec['[jz_beam]'].shape == (100,)
    # energy balance sign: the driver (charge -1, jz_beam < 0) is decelerated: E.j < 0 overall
    assert rec['integrated']['[Ez*jz_beam]'] < 0


@pytest.mark.gpu
def test_cuda_field_insitu_matches_oracle(repo_root, tmp_path):
    import hipace_b200 as hp
    deck = open(os.path.join(repo_root, 'examples', 'blowout_wake_normalized.in')).read()
    ov = {'amr.n_cell': '32 32 40', 'fields.insitu_period': 1,
          'fields.insitu_file_prefix': str(tmp_path / 'f')}
    sim = hp.Simulation(deck, ov)
    sim.evolve(0, 0)
    sim.close()
    got = hp.read_insitu(tmp_path / 'f' / 'reduced_fields.0000.txt')
    ref = O.Simulation(deck, ov)
    ref.evolve(step_end=0)
    want = ref.field_insitu_records[0]
    assert got.shape == (1,) and got.dtype == want.dtype
    scale = max(abs(want['integrated'][nm]) for nm in want['integrated'].dtype.names)
    for nm in want.dtype.names:
        if nm == 'integrated':
            for sub in want[nm].dtype.names:
                assert got[0][nm][sub] == pytest.approx(want[nm][sub], rel=1e-9, abs=1e-12 * scale), sub
        else:
            assert np.allclose(got[0][nm], want[nm], rtol=1e-9, atol=1e-12 * scale), nm


def test_laser_insitu_writer_terms_and_oracle(tmp_path, repo_root):
    import hipace_b200 as hp
    from hipace_b200.build import build_host_check
    rng = np.random.default_rng(13)
    hc = C.CDLL(build_host_check())
    for nx, ny in ((16, 12), (15, 11), (16, 11)):        # 4, 1 and 2 centre cells
        sums = rng.uniform(0.1, 2.0, (8, 5))
        path = tmp_path / f'reduced_laser.{nx}.{ny}.txt'
        hp.insitu_write_laser(path, 0.5, 3, -1e-4, 1e-4, False, 1e-18, nx, ny, sums)
        ours = hp.read_insitu(path)
        dt, rec = O.insitu_laser_record(sums, 0.5, 3, -1e-4, 1e-4, False, 1e-18, nx, ny)
        assert dt == ours.dtype and ours[0].tobytes() == rec.tobytes()
        assert ours[0]['integrated']['max(|a|^2)'] == sums[0].max()
        # terms on the host against the oracle
        geom = O.Geometry(nx, ny, 5, (-3e-5, -2e-5, 0.0), (5e-5, 7e-5, 1e-5), 2, 2)
        env = np.ascontiguousarray(rng.normal(size=(ny, nx)) + 1j * rng.normal(size=(ny, nx)))
        want = O.laser_insitu_sums(env, geom)
        out = np.zeros(8)
        hc.hc_laser_insitu(env.ctypes.data_as(C.c_void_p), nx, ny, C.c_double(geom.dx), C.c_double(geom.dy),
                           C.c_double(geom.pos_offset(0)), C.c_double(geom.pos_offset(1)),
                           out.ctypes.data_as(C.c_void_p))
        assert np.allclose(out, want, rtol=1e-12, atol=1e-25)
    if os.path.isdir(REF_TOOLS):
        sys.path.insert(0, REF_TOOLS)
        import read_insitu_diagnostics as R
        theirs = R.read_file(str(tmp_path / 'reduced_laser.16.12.txt'))
        assert theirs.dtype['axis(a)'].base == np.dtype('<c16')
    # content: the vacuum pulse keeps its energy from step to step (3 steps of the golden's deck)
    deck = open(os.path.join(repo_root, 'examples', 'laser_vacuum_SI.in')).read()
    sim = O.Simulation(deck, {'lasers.solver_type': 'fft', 'max_step': 2, 'lasers.insitu_period': 1,
                              'amr.n_cell': '64 64 50'})
    sim.evolve(step_end=2)
    e = [float(r['integrated']['[|a|^2]']) for r in sim.laser_insitu_records]
    assert len(e) == 3 and abs(e[2] - e[0]) <= 2e-3 * e[0]
