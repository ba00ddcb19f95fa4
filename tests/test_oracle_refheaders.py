"""The NumPy restatement against the REFERENCE ITSELF, for the part of the path that is
header-only in the reference: oracle/ref_headers.cpp compiles ShapeFactors.H,
PushPlasmaParticles.H and DualNumbers.H from /root/reference (in the build container; the built
library travels).  Every shape order 0..3, every derivative type 0..2, the plain and the
dual-number momentum derivative."""
import numpy as np
import pytest

import oracle.hipace_oracle as O
from oracle import refhdr

pytestmark = pytest.mark.skipif(refhdr.lib() is None,
                                reason='neither /root/reference nor a built oracle/_ref/libhipace_refhdr.so')


def _positions():
    rng = np.random.default_rng(7)
    x = rng.uniform(-3.0, 70.0, 50000)
    # cell centres, cell faces and their neighbours in floating point: the branch points
    k = np.arange(-3, 12, dtype=float)
    edge = np.concatenate([k, k + 0.5, np.nextafter(k, 100), np.nextafter(k, -100),
                           np.nextafter(k + 0.5, 100), np.nextafter(k + 0.5, -100)])
    return np.concatenate([x, edge])


@pytest.mark.parametrize('order', [0, 1, 2, 3])
def test_shape_factors_match_reference_header(order):
    x = _positions()
    s_arr, s_single, s_branchless, cell = refhdr.ref_shape(order, x)
    assert (cell > -(1 << 39)).all(), 'the reference variants disagree on the cell'
    s, j0 = O.shape(order, x)
    assert (j0 == cell).all()
    assert np.abs(np.asarray(s) - s_arr).max() <= 4e-16
    assert np.abs(np.asarray(s) - s_single).max() <= 4e-16
    assert np.abs(np.asarray(s) - s_branchless).max() <= 1e-15


@pytest.mark.parametrize('dtype', [0, 1, 2])
@pytest.mark.parametrize('order', [0, 1, 2, 3])
def test_derivative_shape_factors_match_reference_header(dtype, order):
    if dtype == 0 and order == 0:
        pytest.skip('rejected by the reference (Hipace.cpp:52-53)')
    x = _positions()
    s_r, ds_r, cell = refhdr.ref_dshape(dtype, order, x)
    assert (cell > -(1 << 39)).all()
    s, ds, j0 = O.dshape(dtype, order, x)
    assert (j0 == cell).all()
    assert np.abs(s - s_r).max() <= 5e-16
    assert np.abs(ds - ds_r).max() <= 5e-16


def test_momentum_derivative_matches_reference_header():
    rng = np.random.default_rng(11)
    n = 20000
    ux, uy = rng.normal(0, 2, n), rng.normal(0, 2, n)
    psi_inv = 1.0 / rng.uniform(0.05, 3.0, n)
    fields = [rng.normal(0, 1.5, n) for _ in range(6)]          # ExmBy EypBx Ez Bx_c By_c Bz
    for laser in (False, True):
        A = rng.uniform(0, 4, n) if laser else np.zeros(n)
        ADx = rng.normal(0, 1, n) if laser else np.zeros(n)
        ADy = rng.normal(0, 1, n) if laser else np.zeros(n)
        for clight_inv, qmc in ((1.0, -1.0), (1.0 / 299792458.0, -1.758820e11 / 299792458.0)):
            inp = [ux, uy, psi_inv] + fields + [A, ADx, ADy]
            want = refhdr.ref_momentum_push(inp, clight_inv, qmc)
            got = O._momentum_push(ux, uy, psi_inv, *fields, clight_inv, qmc, A, ADx, ADy)
            for g, w in zip(got, want):
                assert np.array_equal(g, w), 'plain-real momentum derivative is not bit-identical'
            eps = [rng.normal(0, 1, n) for _ in range(3)]
            val, ep = refhdr.ref_momentum_push_dual(inp, eps, clight_inv, qmc)
            got_e = O._momentum_push_dual(ux, eps[0], uy, eps[1], psi_inv, eps[2], *fields,
                                          clight_inv, qmc, A, ADx, ADy)
            for g, w in zip(got_e, ep):
                assert np.array_equal(g, w), 'dual-number epsilon part is not bit-identical'
            for g, w in zip(got, val):
                assert np.array_equal(g, w)


@pytest.mark.parametrize('order', [0, 1, 2, 3])
def test_field_gather_and_laser_gather_match_reference_header(order):
    """gather_fields / laser_gather of the oracle against doGatherShapeN / doLaserGatherShapeN of
    the reference's FieldGather.H on random planes (SI-like spacings, non-square grid)"""
    rng = np.random.default_rng(5 + order)
    geom = O.Geometry(24, 40, 8, (-3e-5, -2e-5, 0.0), (5e-5, 7e-5, 1e-5), order, 2)
    g = geom.g
    names = ('Psi', 'Ez', 'Bx', 'By', 'Bz')
    F = {('This', nm): rng.normal(0, 1, (geom.ny + 2 * g, geom.nx + 2 * g)) for nm in names}
    n = 5000
    xp = rng.uniform(geom.lo[0], geom.hi[0], n)
    yp = rng.uniform(geom.lo[1], geom.hi[1], n)
    got = O.gather_fields(xp, yp, F, geom)
    planes = np.stack([F[('This', nm)] for nm in names])
    want = refhdr.ref_gather(order, xp, yp, planes, g, [0, 1, 2, 3, 4], 1.0 / geom.dx, 1.0 / geom.dy,
                             geom.pos_offset(0), geom.pos_offset(1))
    for a, b, scale in zip(got, want, (1 / geom.dx, 1 / geom.dy, 1, 1, 1, 1)):
        assert np.abs(a - b).max() <= 2e-14 * scale * 4
    aabs = rng.uniform(0, 3, (geom.ny + 2 * g, geom.nx + 2 * g))
    # keep away from the last guard ring: the on-the-fly derivative reads one cell further out
    xq = rng.uniform(geom.lo[0] + geom.dx, geom.hi[0] - geom.dx, n)
    yq = rng.uniform(geom.lo[1] + geom.dy, geom.hi[1] - geom.dy, n)
    A, ADx, ADy = O.laser_gather(xq, yq, aabs, geom, True)
    r = refhdr.ref_laser_gather(order, xq, yq, aabs, g, 1.0 / geom.dx, 1.0 / geom.dy,
                                geom.pos_offset(0), geom.pos_offset(1))
    assert np.abs(A - r[0]).max() <= 1e-14 and np.abs(A - r[3]).max() <= 1e-14
    assert np.abs(ADx - r[1]).max() <= 1e-13 / geom.dx
    assert np.abs(ADy - r[2]).max() <= 1e-13 / geom.dy


@pytest.mark.parametrize('monopole', [True, False])
def test_open_boundary_multipoles_match_reference_header(monopole):
    """the complex-moment form of the open-boundary expansion (oracle open_boundary_rhs) against the
    37 tabulated polynomials of the reference's fields/OpenBoundary.H: sources inside the cut-off
    circle, field values on the ring of guard cells"""
    rng = np.random.default_rng(17)
    geom = O.Geometry(40, 32, 4, (-4.0, -3.0, -1.0), (5.0, 3.5, 1.0), 2, 2)
    nx, ny, dx, dy = geom.nx, geom.ny, geom.dx, geom.dy
    rhs = rng.normal(0, 1, (ny, nx))
    got = O.open_boundary_rhs(rhs, geom, monopole) - rhs              # what the boundary adds
    # the same from the reference's functions, assembled as Fields::SetBoundaryCondition does
    off_x = 0.5 * (geom.lo[0] + geom.hi[0] - dx * (nx - 1))
    off_y = 0.5 * (geom.lo[1] + geom.hi[1] - dy * (ny - 1))
    scale = 3.0 / np.hypot(geom.hi[0] - geom.lo[0], geom.hi[1] - geom.lo[1])
    radius = min(abs(geom.lo[0]), abs(geom.hi[0]), abs(geom.lo[1]), abs(geom.hi[1]))
    X, Y = np.meshgrid((np.arange(nx) * dx + off_x) * scale, (np.arange(ny) * dy + off_y) * scale)
    inside = ~(X * X + Y * Y > (0.95 * radius * scale) ** 2)
    xi, yj = np.arange(nx) * dx + off_x, np.arange(ny) * dy + off_y
    edges = [(xi, np.full(nx, -dy + off_y), dy * dy, (0, slice(None))),
             (xi, np.full(nx, ny * dy + off_y), dy * dy, (ny - 1, slice(None))),
             (np.full(ny, -dx + off_x), yj, dx * dx, (slice(None), 0)),
             (np.full(ny, nx * dx + off_x), yj, dx * dx, (slice(None), nx - 1))]
    want = np.zeros_like(rhs)
    for xd, yd, dd, where in edges:
        phi = refhdr.ref_open_boundary(rhs[inside], X[inside], Y[inside], monopole, xd * scale, yd * scale)
        want[where] += -(dx * dy / (4 * np.pi)) * phi / dd
    assert np.abs(want).max() > 0
    assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max()
