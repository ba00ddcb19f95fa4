"""boundary.field = Periodic / fields.poisson_solver = FFTPeriodic: the oracle's restatement of
Fields::EnforcePeriodic (fields/Fields.cpp:1117-1145) and FFTPoissonSolverPeriodic
(fields/fft_poisson_solver/FFTPoissonSolverPeriodic.cpp:69-149) held to what those operations are
defined to be, and the deck surface.  No reference deck, test or golden uses periodic FIELDS (every
example sets boundary.field = Dirichlet): parity of this option is unpinned, the properties below and
the CUDA-vs-oracle tests of tests/test_gpu_periodic.py are what hold it."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('nx,ny', [(32, 32), (30, 24), (33, 27)])
def test_periodic_poisson_on_single_fourier_modes(nx, ny):
    """laplace(phi) = rhs for rhs = one Fourier mode: phi = -rhs / k^2 -- except that the reference
    zeroes inv_k2 on the whole kx = 0 row and ky = 0 column (:84-89), so modes that vary along one
    axis only come out as 0"""
    from oracle.hipace_oracle import poisson_periodic
    dx, dy = 0.3, 0.25
    x = (np.arange(nx) + 0.5) * dx
    y = (np.arange(ny) + 0.5) * dy
    for mx, my in ((1, 1), (2, 3), (3, -2), (nx // 2, 1)):
        kx, ky = 2 * np.pi * mx / (nx * dx), 2 * np.pi * my / (ny * dy)
        rhs = np.cos(kx * x[None, :] + ky * y[:, None] + 0.3)
        phi = poisson_periodic(rhs, dx, dy)
        want = -rhs / (kx * kx + ky * ky)
        assert np.abs(phi - want).max() <= 1e-12 * np.abs(want).max(), (mx, my)
    for mx, my in ((2, 0), (0, 3), (0, 0)):
        kx, ky = 2 * np.pi * mx / (nx * dx), 2 * np.pi * my / (ny * dy)
        rhs = np.cos(kx * x[None, :] + ky * y[:, None] + 0.3)
        assert np.abs(poisson_periodic(rhs, dx, dy)).max() <= 1e-13
    # linear
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal((2, ny, nx))
    lhs = poisson_periodic(a + 2.5 * b, dx, dy)
    assert np.abs(lhs - poisson_periodic(a, dx, dy) - 2.5 * poisson_periodic(b, dx, dy)).max() <= 1e-12


@pytest.mark.parametrize('G', [1, 2, 3])
def test_enforce_periodic_is_sum_and_fill_boundary(G):
    from oracle.hipace_oracle import enforce_periodic
    rng = np.random.default_rng(G)
    ny, nx = 12, 17
    a = rng.standard_normal((ny + 2 * G, nx + 2 * G))
    s = a.copy()
    enforce_periodic([s], G, True)
    # SumBoundary: nothing is lost (the valid box now holds the total), guard cells keep their values,
    # every valid cell holds itself + its images
    assert abs(s[G:-G, G:-G].sum() - a.sum()) <= 1e-12
    m = np.ones_like(a, bool)
    m[G:-G, G:-G] = False
    assert np.array_equal(s[m], a[m])
    for (j, i) in ((0, 0), (0, 5), (ny - 1, nx - 1), (4, nx - 1), (G - 1, G - 1), (G, G)):
        want = 0.0
        for sj in (-ny, 0, ny):
            for si in (-nx, 0, nx):
                jj, ii = j + sj + G, i + si + G
                if 0 <= jj < a.shape[0] and 0 <= ii < a.shape[1]:
                    want += a[jj, ii]
        assert abs(s[j + G, i + G] - want) <= 1e-13
    f = a.copy()
    enforce_periodic([f], G, False)
    assert np.array_equal(f[G:-G, G:-G], a[G:-G, G:-G])
    for jj in range(f.shape[0]):
        for ii in range(f.shape[1]):
            assert f[jj, ii] == a[(jj - G) % ny + G, (ii - G) % nx + G]


def test_periodic_decks_are_accepted_and_described():
    import hipace_b200 as hp
    text = open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()
    hp.deck_check(text, {'boundary.field': 'Periodic'})
    hp.deck_check(text, {'boundary.field': 'Periodic', 'fields.poisson_solver': 'FFTPeriodic'})
    hp.deck_check(text, {'fields.poisson_solver': 'FFTDirichletDirect'})


def test_oracle_periodic_run_differs_from_dirichlet_only_through_the_boundary():
    """a wake in a box so narrow that the sheath reaches the edges: periodic images change the fields;
    with the plasma far from the edges and a Dirichlet Poisson solver the two boundary kinds agree"""
    from oracle.hipace_oracle import Simulation as Oracle
    text = open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()
    ov = {'amr.n_cell': '32 32 40', 'plasma.ppc': '1 1'}
    runs = {}
    for bf in ('Dirichlet', 'Periodic'):
        o = Oracle(text, dict(ov, **{'boundary.field': bf}))
        runs[bf] = o.evolve(30)
        assert all(np.isfinite(v) for v in runs[bf].values())
    assert any(abs(runs['Periodic'][k] - w) > 1e-6 * abs(w) for k, w in runs['Dirichlet'].items() if w)
    # a plasma column that stays away from the guard cells: nothing is deposited there, SumBoundary adds
    # zeros, and FillBoundary only changes guard cells nobody gathers from
    ov2 = dict(ov, **{'plasma.radius': 4.0, 'boundary.particle': 'Absorbing'})
    a = Oracle(text, dict(ov2, **{'boundary.field': 'Dirichlet'})).evolve(12)
    b = Oracle(text, dict(ov2, **{'boundary.field': 'Periodic'})).evolve(12)
    for k, w in a.items():
        if k in ('ExmBy', 'EypBx'):
            continue            # computed one guard ring out, from the (filled) guard cells of Psi
        assert abs(b[k] - w) <= 1e-9 * abs(w) + 1e-30, (k, b[k], w)
