"""GPU parity: the CUDA slice loop (through the C-ABI of include/hpb200.h) against the oracle and
against the reference's own golden checksums.

Tolerances: the reference accepts rtol 2e-5 (blowout) / 1e-7 (linear wake) between
its CPU and CUDA builds because of atomic summation order; we hold ourselves to RTOL_SUM = 1e-9
on sum|Q| (the reference's default checksum tolerance) and RTOL_CELL = 1e-9 of the field's
max-norm per cell.  Particle counts and validity flags are bit-exact.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
RTOL_SUM = 1e-9
RTOL_CELL = 1e-9
FLOOR = 1e-3      # round-off of the O(1) plasma/ion densities puts ~1e-15 absolute noise on every field


def _deck(repo_root, name):
    return open(os.path.join(repo_root, 'examples', name)).read()


@pytest.mark.parametrize('case', ['linear_wake.normalized.1Rank', 'blowout_wake_explicit.2Rank',
                                  'laser_blowout_wake_explicit.SI.1Rank',
                                  'laser_blowout_wake_explicit.1Rank', 'linear_wake.SI.1Rank'])
def test_slice_loop_matches_reference_golden(case, repo_root):
    import hipace_b200 as hp
    meta = json.load(open(os.path.join(GOLD, case + '.json')))
    ov = dict(meta['overrides'])
    ov.pop('max_step', None)          # dt = 0: every step repeats step 0 (tests/blowout_wake_explicit.2Rank.sh)
    sim = hp.Simulation(open(os.path.join(repo_root, meta['deck'])).read(), ov)
    cs = sim.evolve()
    gold = meta['checksums']
    for name, want in gold['lev=0'].items():
        got = cs[name]
        assert abs(got - want) <= RTOL_SUM * abs(want) + 1e-40, (name, got, want)
    bc = sim.beam_checksums() if 'beam' in gold else {}
    for name, want in gold.get('beam', {}).items():
        if name in bc:
            assert abs(bc[name] - want) <= RTOL_SUM * abs(want) + 1e-40, (name, bc[name], want)
    assert sim.stats()['n_qsa_violation'] == 0
    sim.close()


@pytest.mark.parametrize('fuse', [0, 1])
@pytest.mark.parametrize('deck,ov,nsl', [
    ('blowout_wake_normalized.in', {}, 60),
    ('blowout_wake_normalized.in', {'amr.n_cell': '63 63 100', 'plasma.ppc': '2 2'}, 40),
    ('linear_wake_normalized.in', {'amr.n_cell': '48 40 100'}, 50),
])
def test_slice_by_slice_fields_and_particles(deck, ov, nsl, fuse, repo_root):
    """Every field component per cell, and the full particle state, after each of the first
    slices (the head of the box, through the beam and into the blow-out).

    fuse = 0: the reference's call order (InitializeSlices ... ShiftSlices), every component is
    comparable after the call.  fuse = 1 (the default driver): the end of a slice has already
    shifted / re-initialised chi, jz_beam, rhomjz, jx, jy for the next slice, deposited the pushed
    plasma into them and (on the side stream) seeded the next slice's Sx, Sy from the beam, so
    only the solved fields are compared per slice -- the sources are covered through the fields
    solved from them (Bx, By of the same slice, everything of the following one) and the final
    particle state."""
    import hipace_b200 as hp
    from oracle.hipace_oracle import Simulation as Oracle
    text = _deck(repo_root, deck)
    ref = Oracle(text, ov)
    sim = hp.Simulation(text, ov)
    sim.set_option('fuse', fuse)
    ref.begin_step()
    sim.begin_step(0)
    names = ['chi', 'Sy', 'Sx', 'ExmBy', 'EypBx', 'Ez', 'Bx', 'By', 'Bz', 'Psi', 'jx_beam',
             'jy_beam', 'jz_beam', 'jx', 'jy', 'rhomjz']
    reset_by_fusion = ('chi', 'jz_beam', 'rhomjz', 'Sy', 'Sx')      # Sx, Sy: already the next slice's beam seed
    nz = ref.geom.nz
    snap = {}
    ref.slice_hook = lambda s, isl, stage: snap.update(
        {n: s.T(n).copy() for n in names}) if stage == 'fields' else None
    for isl in range(nz - 1, nz - 1 - nsl, -1):
        ref.solve_one_slice(isl)
        # run the CUDA slice up to the same point: fields are final before the push, and the
        # push does not modify them, so compare after the whole slice except the shifted comps
        sim.solve_one_slice(isl)
        for n in names:
            if n in ('jx', 'jy', 'jx_beam', 'jy_beam') or (fuse and n in reset_by_fusion):
                continue        # rotated by ShiftSlices at the end of the slice
            a, b = sim.field(n), snap[n]
            # ahead of the beam every field is round-off noise around 0 (plasma and ion
            # background cancel): FLOOR keeps the comparison meaningful there
            scale = max(np.abs(b).max(), FLOOR)
            err = np.abs(a - b).max() / scale
            assert err <= RTOL_CELL, (isl, n, err)
        for n, want in (('jx', ref.F[('This', 'jx')]), ('jy', ref.F[('This', 'jy')]),
                        ('jx_beam', ref.F[('This', 'jx_beam')]), ('jy_beam', ref.F[('Previous', 'jy_beam')])):
            if fuse and n in ('jx', 'jy'):
                continue        # already hold the next slice's plasma current on top of the beam's
            got = sim.field(n, 'Previous' if n == 'jy_beam' else 'This')
            err = np.abs(got - want).max() / max(np.abs(want).max(), FLOOR)
            assert err <= RTOL_CELL, (isl, n, err, np.abs(got).max(), np.abs(want).max())
    p = sim.plasma()
    o = ref.plasmas[0]
    assert p['x'].size == o.x.size
    assert np.array_equal(p['valid'], o.valid)                      # bit-exact
    v = o.valid
    for nm, ov_ in (('x', o.x), ('y', o.y), ('ux', o.ux), ('uy', o.uy), ('psi', o.psi),
                    ('x_prev', o.x_prev), ('ux_half_step', o.ux_half), ('psi_half_step', o.psi_half),
                    ('w', o.w)):
        scale = max(np.abs(ov_[v]).max(), 1e-300)
        err = np.abs(p[nm][v] - ov_[v]).max() / scale
        assert err <= 1e-9, (nm, err)
    sim.close()


@pytest.mark.parametrize('fuse', [0, 1])
def test_two_mobile_species_ion_motion(fuse, repo_root):
    """two mobile species (electrons + light ions, no neutralising background; the
    structure of the reference's inputs_ion_motion_SI with a deterministic beam) -- per-species
    charge / mass through every particle kernel, both driver orders, vs the oracle: field
    checksums of 70 slices and the full state of both species."""
    import hipace_b200 as hp
    from oracle.hipace_oracle import Simulation as Oracle
    text = _deck(repo_root, 'ion_motion_normalized.in')
    nsl = 70
    ref = Oracle(text, {})
    want = ref.evolve(nsl)
    sim = hp.Simulation(text, {})
    sim.set_option('fuse', fuse)
    got = sim.evolve(0, 0, nsl)
    for k, w in want.items():
        assert abs(got[k] - w) <= RTOL_SUM * abs(w) + 1e-30, (k, got[k], w)
    assert sim.stats()['n_qsa_violation'] == 0
    for isp, o in enumerate(ref.plasmas):
        p = sim.plasma(isp)
        assert np.array_equal(p['valid'], o.valid)
        v = o.valid
        for nm, ov_ in (('x', o.x), ('y', o.y), ('ux', o.ux), ('uy', o.uy), ('psi', o.psi)):
            scale = max(np.abs(ov_[v]).max(), 1e-300)
            assert np.abs(p[nm][v] - ov_[v]).max() / scale <= 1e-9, (isp, nm)
    sim.close()


def test_production_shape_ppc9_two_species(repo_root):
    """BASELINE configs[4] in miniature: ppc 9 (3 x 3) for BOTH mobile species (electrons + ions, ion motion
    on, no neutralising background) on 96 x 96 -- the shape of the reference's production deck -- against
    the oracle: field checksums of 30 slices, validity bit-exact, the state of both species to 1e-9"""
    import hipace_b200 as hp
    from oracle.hipace_oracle import Simulation as Oracle
    text = _deck(repo_root, 'ion_motion_normalized.in')
    ov = {'amr.n_cell': '96 96 100', 'elec.ppc': '3 3', 'ions.ppc': '3 3'}
    nsl = 30
    ref = Oracle(text, ov)
    want = ref.evolve(nsl)
    sim = hp.Simulation(text, ov)
    got = sim.evolve(0, 0, nsl)
    for k, w in want.items():
        assert abs(got[k] - w) <= RTOL_SUM * abs(w) + 1e-30, (k, got[k], w)
    assert sim.stats()['n_qsa_violation'] == 0
    for isp, o in enumerate(ref.plasmas):
        p = sim.plasma(isp)
        assert p['x'].size == 9 * 96 * 96
        assert np.array_equal(p['valid'], o.valid)
        v = o.valid
        for nm, ov_ in (('x', o.x), ('y', o.y), ('ux', o.ux), ('uy', o.uy), ('psi', o.psi)):
            scale = max(np.abs(ov_[v]).max(), 1e-300)
            assert np.abs(p[nm][v] - ov_[v]).max() / scale <= 1e-9, (isp, nm)
    sim.close()


def test_full_size_driver_orders_agree(repo_root):
    """BASELINE configs[2] transverse size (1024 x 1024, ppc 4; the oracle would need minutes per
    slice here): the reference call order, the fused order and the fused order with the side
    stream must give the same physics -- field checksums of 48 slices through the beam head to
    1e-9, particle validity bit-exact, particle state to 1e-9, identical multigrid V-cycle counts;
    plus the invariant the domain offers: no particle is lost or invalidated (periodic boundary,
    no QSA violation)."""
    import hipace_b200 as hp
    text = _deck(repo_root, 'blowout_wake_normalized.in')
    ov = {'amr.n_cell': '1024 1024 1024', 'plasma.ppc': '2 2'}
    nsl = 48
    runs = {}
    for name, opts in (('reference order', {'fuse': 0}), ('fused', {'fuse': 1, 'side_stream': 0}),
                       ('fused + side stream', {'fuse': 1, 'side_stream': 1})):
        sim = hp.Simulation(text, ov)
        for k, v in opts.items():
            sim.set_option(k, v)
        cs = sim.evolve(0, 0, nsl)
        p = sim.plasma()
        st = sim.stats()
        runs[name] = (cs, p, st)
        sim.close()
    cs0, p0, st0 = runs['reference order']
    assert p0['x'].size == 4 * 1024 * 1024 and p0['valid'].all() and st0['n_qsa_violation'] == 0
    for name in ('fused', 'fused + side stream'):
        cs, p, st = runs[name]
        for k, w in cs0.items():
            assert abs(cs[k] - w) <= RTOL_SUM * abs(w) + 1e-30, (name, k, cs[k], w)
        assert np.array_equal(p['valid'], p0['valid']), name
        for nm in ('x', 'y', 'ux', 'uy', 'psi', 'w'):
            scale = max(np.abs(p0[nm]).max(), 1e-300)
            assert np.abs(p[nm] - p0[nm]).max() / scale <= 1e-9, (name, nm)
        assert st['n_mg_vcycles'] == st0['n_mg_vcycles'], name


@pytest.mark.parametrize('n_cell,nsl', [('1024 1024 1024', 48), ('1023 1023 1024', 48), ('256 256 512', 160)],
                         ids=['configs2_1024', 'recommended_1023', 'configs1_256x512'])
def test_full_size_matches_cport(n_cell, nsl, repo_root):
    """The CUDA slice loop against the CPU restatement (oracle/cport: the C/OpenMP companion of the
    NumPy oracle, itself pinned to the reference goldens in tests/test_cport.py) AT the benchmark
    sizes: BASELINE configs[2] (1024 x 1024, ppc 4), the reference's recommended 2^n - 1 grid
    (1023 x 1023) and configs[1] (256 x 256 x 512, ppc 4) -- nsl slices from the head of the box
    through the beam head.  Field checksums to 1e-9, particle validity bit-exact, particle state to
    1e-9, the same multigrid V-cycle count on every slice.  (Bz vanishes by symmetry for this deck:
    its checksum is the sum of round-off sized values and is held to 1e-12 of the largest checksum,
    like the other symmetry-zero quantities of this suite.)"""
    import hipace_b200 as hp
    from oracle import cport
    text = _deck(repo_root, 'blowout_wake_normalized.in')
    ov = {'amr.n_cell': n_cell, 'plasma.ppc': '2 2'}
    cport.set_threads(cport.physical_cores())
    ref = cport.Simulation(text, ov)
    want = ref.evolve(nsl)
    sim = hp.Simulation(text, ov)
    got = sim.evolve(0, 0, nsl)
    atol = 1e-12 * max(abs(w) for w in want.values())
    for k, w in want.items():
        assert abs(got[k] - w) <= RTOL_SUM * abs(w) + atol, (k, got[k], w)
    # V-cycles per slice.  Ahead of the beam the right-hand side is zero up to the round-off of the
    # deposition sums (exactly zero on some grids and summation orders, 1e-20 on others): hpmg's
    # relative stopping rule then iterates on noise or not at all, so only slices where the
    # reference solver has a physical source (>= 1 V-cycle) are compared
    got_it, want_it = sim.mg_iters(), list(ref.mg_cycles)
    assert len(got_it) == len(want_it)
    assert [g for g, w in zip(got_it, want_it) if w > 0] == [w for w in want_it if w > 0], 'V-cycles per slice'
    assert max(ref.mg_cycles) >= 2, 'the sample must reach the beam (more than the vacuum V-cycle)'
    o, p = ref.plasmas[0], sim.plasma()
    assert np.array_equal(p['valid'], o.valid)
    for nm in ('x', 'y', 'ux', 'uy', 'psi', 'w'):
        want_a = getattr(o, nm)
        scale = max(np.abs(want_a).max(), 1e-300)
        assert np.abs(p[nm] - want_a).max() / scale <= 1e-9, nm
    sim.close()


def test_reference_gpu_arm_matches_the_product_solver(repo_root):
    """bench.py --impl cufft_ref / naive: the reference's DirichletFast sequence on cuFFT (ref_gpu_arm.cu)
    and the generic one-thread-per-particle kernels in the reference call order must reproduce the
    reference golden like the product path does -- otherwise the bar they set would be meaningless"""
    import hipace_b200 as hp
    meta = json.load(open(os.path.join(GOLD, 'blowout_wake_explicit.2Rank.json')))
    ov = dict(meta['overrides'])
    ov.pop('max_step', None)
    text = open(os.path.join(repo_root, meta['deck'])).read()
    for opts in ({'poisson_impl': 1}, {'poisson_impl': 1, 'generic_order_kernels': 1, 'fuse': 0, 'side_stream': 0}):
        sim = hp.Simulation(text, ov)
        for k, v in opts.items():
            sim.set_option(k, v)
        cs = sim.evolve()
        for name, want in meta['checksums']['lev=0'].items():
            assert abs(cs[name] - want) <= RTOL_SUM * abs(want) + 1e-40, (opts, name, cs[name], want)
        sim.close()


@pytest.mark.parametrize('idx_type', ['0 0', '1 1'])
def test_plasma_reorder_invariance(idx_type, repo_root):
    """plasmas.reorder_period = 4: the reference re-runs its blowout_wake_explicit golden with the plasma
    sorted by cell every 4 slices and expects the same checksums
    (/root/reference/tests/blowout_wake_explicit.2Rank.sh:45-58).  Here: the reference golden itself with
    the sort on, plus -- against the un-sorted run -- the same particle SET (sorted values bit-exact at the
    sort, states equal to 1e-9 after 100 slices, multiset comparison) and V-cycle counts."""
    import hipace_b200 as hp
    meta = json.load(open(os.path.join(GOLD, 'blowout_wake_explicit.2Rank.json')))
    text = open(os.path.join(repo_root, meta['deck'])).read()
    ov = dict(meta['overrides'])
    ov.pop('max_step', None)
    ov['plasma.ppc'] = '2 2'
    runs = {}
    for period in (0, 4):
        sim = hp.Simulation(text, dict(ov, **{'plasmas.reorder_period': period, 'plasmas.reorder_idx_type': idx_type}))
        cs = sim.evolve()
        runs[period] = (cs, sim.plasma(), sim.stats(), sim.mg_iters())
        sim.close()
    cs0, p0, st0, it0 = runs[0]
    cs4, p4, st4, it4 = runs[4]
    assert st0['n_reorders'] == 0 and st4['n_reorders'] == 25          # 100 slices, every 4th
    for k, w in cs0.items():
        assert abs(cs4[k] - w) <= RTOL_SUM * abs(w) + 1e-30, (k, cs4[k], w)
    assert it4 == it0
    assert p4['valid'].sum() == p0['valid'].sum()
    # the same particles in another order: the order statistics of every component agree
    for nm in ('x', 'y', 'ux', 'uy', 'psi', 'w'):
        a, b = np.sort(p0[nm][p0['valid']]), np.sort(p4[nm][p4['valid']])
        scale = max(np.abs(a).max(), 1e-300)
        assert np.abs(a - b).max() / scale <= 1e-8, nm


def test_reference_golden_with_reordering(repo_root):
    """the reference's own check: the blowout_wake_explicit golden must come out with reorder_period = 4"""
    import hipace_b200 as hp
    meta = json.load(open(os.path.join(GOLD, 'blowout_wake_explicit.2Rank.json')))
    ov = dict(meta['overrides'])
    ov.pop('max_step', None)
    ov['plasmas.reorder_period'] = 4
    sim = hp.Simulation(open(os.path.join(repo_root, meta['deck'])).read(), ov)
    cs = sim.evolve()
    for name, want in meta['checksums']['lev=0'].items():
        assert abs(cs[name] - want) <= RTOL_SUM * abs(want) + 1e-40, (name, cs[name], want)
    assert sim.stats()['n_reorders'] == 25
    sim.close()


def test_si_units_deck_matches_oracle(repo_root):
    """hipace.normalized_units = 0: the SI constants, invvol = 1/(dx dy dz) and the dx dy dz / ppc
    weights through every kernel (the reference's examples/blowout_wake/inputs_SI restated).  The
    oracle's SI path is pinned by the reference's laser_blowout_wake_explicit.SI.1Rank golden."""
    import hipace_b200 as hp
    from oracle.hipace_oracle import Simulation as Oracle
    text = _deck(repo_root, 'blowout_wake_SI.in')
    nsl = 45
    want = Oracle(text, {}).evolve(nsl)
    for fuse in (0, 1):
        sim = hp.Simulation(text, {})
        sim.set_option('fuse', fuse)
        got = sim.evolve(0, 0, nsl)
        for k, w in want.items():
            assert abs(got[k] - w) <= RTOL_SUM * abs(w) + 1e-300, (fuse, k, got[k], w)
        sim.close()


def test_plasma_init_is_bit_exact(repo_root):
    """particle count, order and positions right after InitParticles"""
    import hipace_b200 as hp
    from oracle.hipace_oracle import Simulation as Oracle
    text = _deck(repo_root, 'blowout_wake_normalized.in')
    ov = {'amr.n_cell': '50 38 20', 'plasma.ppc': '3 2', 'plasma.radius': '6.5',
          'plasma.density(x,y,z)': '1. + 0.1*x'}
    ref = Oracle(text, ov)
    sim = hp.Simulation(text, ov)
    ref.begin_step()
    sim.begin_step(0)
    p, o = sim.plasma(), ref.plasmas[0]
    assert p['x'].size == o.x.size
    assert np.array_equal(p['x'], o.x) and np.array_equal(p['y'], o.y)
    assert np.allclose(p['w'], o.w, rtol=1e-15, atol=0)
    a = sim.field('rhomjz', 'RhomJzIons')
    b = ref.F[('RhomJzIons', 'rhomjz')]
    assert np.abs(a - b).max() <= 1e-13 * np.abs(b).max()
    sim.close()


@pytest.mark.parametrize('nx,ny', [(64, 64), (63, 63), (100, 36), (255, 129), (1024, 1024),
                                   (1023, 1023),
                                   # nx + 1 with a large prime factor: the Bluestein path of fft_smem.cuh
                                   (256, 256), (2048, 48), (682, 40)])
def test_poisson_solve_matches_dst_oracle(nx, ny):
    """FFTPoissonSolverDirichlet: lhs = DST2D(DST2D(rhs) * eigenvalues) (oracle) vs the CUDA
    row-DST + tridiagonal formulation, incl. non power-of-two N = nx + 1 (1025 = 5*5*41)."""
    import torch
    import hipace_b200 as hp
    from oracle.hipace_oracle import poisson_dirichlet, poisson_eigenvalues
    rng = np.random.default_rng(nx * 1000 + ny)
    dx, dy = 16. / nx, 12. / ny
    rhs = rng.standard_normal((3, ny, nx))
    ctx = hp.Context(nx, ny, dx, dy, 0.1, -8 + dx / 2, -6 + dy / 2)
    g = hp.NGUARD
    sl_t = torch.full((4, ny + 2 * g, nx + 2 * g), 7.0, dtype=torch.float64, device='cuda')
    ctx.poisson_solve(torch.from_numpy(rhs).cuda(), ctx.slice_view(sl_t), [0, 2, 3])
    torch.cuda.synchronize()
    out = sl_t.cpu().numpy()
    eig = poisson_eigenvalues(nx, ny, dx, dy)
    for b, c in enumerate([0, 2, 3]):
        want = poisson_dirichlet(rhs[b], eig)
        err = np.abs(out[c, g:-g, g:-g] - want).max() / np.abs(want).max()
        assert err <= 1e-11, (b, err)
    # guard cells and untouched components keep their values
    assert (out[1] == 7.0).all()
    assert (out[0, :g] == 7.0).all() and (out[0, :, :g] == 7.0).all()
    ctx.close()


@pytest.mark.parametrize('p,m', [(7, 8), (11, 6), (13, 4), (17, 4), (19, 4), (23, 4), (29, 2), (31, 2), (37, 2),
                                 (41, 2), (43, 2), (47, 2), (53, 2), (59, 2), (61, 2), (41, 25), (7, 49)])
def test_poisson_prime_stage_on_tensor_cores(p, m):
    """the odd-prime FFT stage (fft_smem.cuh: P = C U, Q = S V as mma.m8n8k4.f64 tiles) for EVERY prime the
    generic stage serves (7 .. 61: all paddings of the (h+1) x h coefficient matrices to 8 x 4 tiles, all
    tile counts of the 2 Nr columns), row length nx + 1 = p m, against the DST oracle"""
    import torch
    import hipace_b200 as hp
    from oracle.hipace_oracle import poisson_dirichlet, poisson_eigenvalues
    nx, ny = p * m - 1, (15 if (p * m - 1) % 2 else 16)
    rng = np.random.default_rng(p * 100 + m)
    dx, dy = 16. / nx, 12. / ny
    rhs = rng.standard_normal((3, ny, nx))
    ctx = hp.Context(nx, ny, dx, dy, 0.1, -8 + dx / 2, -6 + dy / 2)
    g = hp.NGUARD
    sl_t = torch.zeros((3, ny + 2 * g, nx + 2 * g), dtype=torch.float64, device='cuda')
    ctx.poisson_solve(torch.from_numpy(rhs).cuda(), ctx.slice_view(sl_t), [0, 1, 2])
    torch.cuda.synchronize()
    out = sl_t.cpu().numpy()
    eig = poisson_eigenvalues(nx, ny, dx, dy)
    for b in range(3):
        want = poisson_dirichlet(rhs[b], eig)
        err = np.abs(out[b, g:-g, g:-g] - want).max() / np.abs(want).max()
        assert err <= 1e-11, (p, m, b, err)
    ctx.close()


@pytest.mark.parametrize('nx,ny', [(64, 64), (63, 63), (96, 64), (127, 63), (256, 256)])
def test_mg_solve1_matches_hpmg_oracle(nx, ny):
    """hpmg solve1: same V-cycle count and the same values as the oracle restatement."""
    import torch
    import hipace_b200 as hp
    from oracle.hipace_oracle import MultiGrid1
    rng = np.random.default_rng(nx + 7 * ny)
    dx, dy = 16. / nx, 16. / ny
    g = hp.NGUARD
    sl = np.zeros((5, ny + 2 * g, nx + 2 * g))
    v = (slice(g, -g), slice(g, -g))
    yy, xx = np.meshgrid(np.linspace(-1, 1, ny), np.linspace(-1, 1, nx), indexing='ij')
    sl[0][v] = 1.0 + 0.5 * np.exp(-4 * (xx ** 2 + yy ** 2)) + 0.05 * rng.random((ny, nx))   # chi
    sl[1][v] = np.exp(-8 * ((xx - .2) ** 2 + yy ** 2)) + 0.01 * rng.standard_normal((ny, nx))
    sl[2][v] = xx * np.exp(-6 * (xx ** 2 + (yy + .1) ** 2))
    sl[3][v] = 0.01 * rng.standard_normal((ny, nx))     # initial guess
    sl[4][v] = 0.0
    mg = MultiGrid1(dx, dy, nx, ny)
    sol = np.stack([sl[3][v], sl[4][v]]).copy()
    mg.solve1(sol, np.stack([sl[1][v], sl[2][v]]), sl[0][v].copy())
    ctx = hp.Context(nx, ny, dx, dy, 0.1, -8 + dx / 2, -8 + dy / 2)
    t = torch.from_numpy(sl).cuda()
    iters = ctx.mg_solve1(ctx.slice_view(t), 3, 1, 0)
    torch.cuda.synchronize()
    out = t.cpu().numpy()
    assert iters == mg.n_vcycles_last
    for k in range(2):
        err = np.abs(out[3 + k][v] - sol[k]).max() / np.abs(sol).max()
        assert err <= 1e-12, (k, err)
    ctx.close()


@pytest.mark.parametrize('nx,ny', [(64, 64), (63, 63), (96, 64), (127, 63), (128, 128)])
def test_mg_solve2_matches_hpmg_oracle(nx, ny):
    """hpmg solve2 (system type 2, the laser envelope's complex Helmholtz system): same V-cycle count and
    the same values as the oracle's MultiGrid2 (gs2 / residual2r / residual2i restated)."""
    import torch
    import hipace_b200 as hp
    from oracle.hipace_oracle import MultiGrid2
    rng = np.random.default_rng(3 * nx + ny)
    dx, dy = 16. / nx, 16. / ny
    yy, xx = np.meshgrid(np.linspace(-1, 1, ny), np.linspace(-1, 1, nx), indexing='ij')
    acf_r = 40.0 + 5.0 * np.exp(-4 * (xx ** 2 + yy ** 2)) + 0.5 * rng.random((ny, nx))
    acf_i = -300.0
    rhs = np.stack([np.exp(-8 * ((xx - .2) ** 2 + yy ** 2)) + 0.01 * rng.standard_normal((ny, nx)),
                    xx * np.exp(-6 * (xx ** 2 + (yy + .1) ** 2))])
    guess = 0.001 * rng.standard_normal((2, ny, nx))
    mg = MultiGrid2(dx, dy, nx, ny)
    want = guess.copy()
    mg.solve2(want, rhs, acf_r, acf_i, 1e-4, 0.0, 200)
    ctx = hp.Context(nx, ny, dx, dy, 0.1, -8 + dx / 2, -8 + dy / 2)
    t_sol, t_rhs, t_acf = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (guess, rhs, acf_r))
    iters = ctx.mg_solve2(t_sol, t_rhs, t_acf, acf_i, 1e-4, 0.0)
    torch.cuda.synchronize()
    got = t_sol.cpu().numpy()
    assert iters == mg.n_vcycles_last and iters >= 1
    assert np.abs(got - want).max() / np.abs(want).max() <= 1e-12
    ctx.close()


def test_command_line_driver_reproduces_golden(repo_root):
    """the C++ driver binary end to end: deck file + key=value overrides in, checksums out"""
    import subprocess
    meta = json.load(open(os.path.join(GOLD, 'linear_wake.normalized.1Rank.json')))
    exe = os.path.join(repo_root, 'hipace_b200', 'bin', 'hpb200_run')
    ov = [f'{k}={v}' for k, v in meta['overrides'].items() if k != 'max_step']
    p = subprocess.run([exe, os.path.join(repo_root, meta['deck'])] + ov, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr + p.stdout
    got = {}
    for line in p.stdout.splitlines():
        t = line.split()
        if len(t) == 2 and line.startswith('  '):
            got[t[0]] = float(t[1])
    for name, want in meta['checksums']['lev=0'].items():
        assert abs(got[name] - want) <= RTOL_SUM * abs(want) + 1e-40, (name, got[name], want)
