"""The arithmetic of the CUDA particle kernels, RUN ON THE CPU.

shapes.cuh / push_math.cuh / generic_order.cuh are __host__ __device__: tests/host_check.cu calls the
very functions the kernels call (per-particle bodies of deposit, beam deposit, explicit deposition,
gather, push; a plain += in place of the fp64 RED) and this file holds them to
  * the reference's own headers (oracle/ref_headers.cpp: ShapeFactors.H, FieldGather.H,
    PushPlasmaParticles.H + DualNumbers.H compiled from /root/reference), and
  * the NumPy oracle (whole-array deposit / explicit deposition / advance),
for every deposition order 0..3 and derivative type 0..2, with and without a laser.
No GPU is involved: this is what can be known about the device code before a GPU run."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle.hipace_oracle as O
from oracle import refhdr
from hipace_b200.build import build_host_check


class HcGrid(C.Structure):
    _fields_ = [('nx_tot', C.c_int), ('ny_tot', C.c_int), ('g', C.c_int), ('x_off', C.c_double),
                ('y_off', C.c_double), ('dx_inv', C.c_double), ('dy_inv', C.c_double)]


class DepositPar(C.Structure):
    _fields_ = [('clightinv', C.c_double), ('charge_invvol', C.c_double),
                ('charge_mu0_mass_ratio', C.c_double), ('max_qsa', C.c_double),
                ('laser_norm', C.c_double), ('c_aabs', C.c_int), ('clight', C.c_double)]


class ExplicitPar(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('c_sy', 'c_sx', 'c_bz', 'c_ez', 'c_exmby', 'c_eypbx', 'c_aabs')] + \
               [(k, C.c_double) for k in ('clight', 'clight_inv', 'charge_invvol_mu0', 'q_mass_ratio', 'laser_fac')]


class PushPar(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('c_psi', 'c_ez', 'c_bx', 'c_by', 'c_bz', 'c_aabs')] + \
               [(k, C.c_double) for k in ('clight', 'qmc', 'dz')] + \
               [(k, C.c_int) for k in ('n_subcycles', 'temp_slice', 'bc')] + \
               [(k, C.c_double) for k in ('lox', 'loy', 'hix', 'hiy', 'laser_norm')]


@pytest.fixture(scope='module')
def hc():
    L = C.CDLL(build_host_check())
    L.hc_deposit_current.restype = C.c_long
    return L


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


def _ptrs(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def _positions():
    rng = np.random.default_rng(7)
    x = rng.uniform(-3.0, 70.0, 20000)
    k = np.arange(-3, 12, dtype=float)
    edge = np.concatenate([k, k + 0.5, np.nextafter(k, 100), np.nextafter(k, -100),
                           np.nextafter(k + 0.5, 100), np.nextafter(k + 0.5, -100)])
    return np.ascontiguousarray(np.concatenate([x, edge]))


needs_ref = pytest.mark.skipif(refhdr.lib() is None, reason='no built oracle/_ref/libhipace_refhdr.so')


@needs_ref
@pytest.mark.parametrize('order', [0, 1, 2, 3])
def test_device_shapes_match_reference_header(hc, order):
    x = _positions()
    n = x.size
    s = np.zeros((order + 1, n)); cell = np.zeros(n, dtype=np.int64)
    assert hc.hc_shape(order, C.c_long(n), _dp(x), _dp(s), _dp(cell)) == 0
    want, _, _, wcell = refhdr.ref_shape(order, x)
    assert (cell == wcell).all()
    assert np.abs(s - want).max() <= 2e-15
    for dtype in (0, 1, 2):
        if order == 0 and dtype == 0:
            continue
        m = order + dtype + 1
        s, ds = np.zeros((m, n)), np.zeros((m, n))
        assert hc.hc_dshape(dtype, order, C.c_long(n), _dp(x), _dp(s), _dp(ds), _dp(cell)) == 0
        ws, wds, wcell = refhdr.ref_dshape(dtype, order, x)
        assert (cell == wcell).all(), (order, dtype)
        assert np.abs(s - ws).max() <= 2e-15, (order, dtype)
        assert np.abs(ds - wds).max() <= 2e-15, (order, dtype)


@needs_ref
def test_device_momentum_push_is_bit_identical_to_reference_header(hc):
    rng = np.random.default_rng(3)
    n = 20000
    base = [rng.normal(0, 2, n), rng.normal(0, 2, n), 1.0 / rng.uniform(0.05, 3.0, n)] + \
           [rng.normal(0, 1.5, n) for _ in range(6)]
    eps = [rng.normal(0, 1, n) for _ in range(3)]
    for laser in (0, 1):
        las = [rng.uniform(0, 4, n), rng.normal(0, 1, n), rng.normal(0, 1, n)] if laser else [np.zeros(n)] * 3
        inp = [np.ascontiguousarray(a) for a in base + las]
        for clight_inv, qmc in ((1.0, -1.0), (1.0 / 299792458.0, -586.6792)):
            out = np.zeros((9, n))
            hc.hc_momentum_push(C.c_long(n), _ptrs(inp), _ptrs(eps), laser, C.c_double(clight_inv),
                                C.c_double(qmc), _dp(out))
            want = refhdr.ref_momentum_push(inp, clight_inv, qmc)
            val, ep = refhdr.ref_momentum_push_dual(inp, eps, clight_inv, qmc)
            assert np.array_equal(out[0:3], want)
            assert np.array_equal(out[3:6], val)
            assert np.array_equal(out[6:9], ep)


def _setup(order, dtype, laser, seed, si=False):
    """random fields and a stirred plasma on a small non-square grid"""
    rng = np.random.default_rng(seed)
    scale = 1e-5 if si else 1.0
    geom = O.Geometry(24, 40, 8, (-3 * scale, -2 * scale, 0.0), (5 * scale, 7 * scale, 1 * scale), order, dtype)
    pc = O.PhysConst.make(not si)
    g = geom.g
    names = ['chi', 'Sy', 'Sx', 'ExmBy', 'EypBx', 'Ez', 'Bx', 'By', 'Bz', 'Psi', 'jx', 'jy', 'rhomjz', 'rho', 'aabs', 'jz']
    fmag = 1e9 if si else 1.0
    F = {('This', nm): np.zeros((geom.ny + 2 * g, geom.nx + 2 * g)) for nm in names}
    for nm in ('ExmBy', 'EypBx', 'Ez', 'Psi'):
        F[('This', nm)][...] = rng.normal(0, fmag * (scale if nm == 'Psi' else 1.0), F[('This', nm)].shape)
    for nm in ('Bx', 'By', 'Bz'):
        F[('This', nm)][...] = rng.normal(0, fmag / pc.c, F[('This', nm)].shape)
    F[('This', 'aabs')][...] = rng.uniform(0, 2, F[('This', 'aabs')].shape)
    n = 3000
    charge, mass = (-pc.q_e, pc.m_e)
    pl = O.Plasma('p', charge, mass, (1, 1), None)
    pl.x = rng.uniform(geom.lo[0] + geom.dx, geom.hi[0] - geom.dx, n)
    pl.y = rng.uniform(geom.lo[1] + geom.dy, geom.hi[1] - geom.dy, n)
    pl.w = rng.uniform(0.5, 1.5, n) * (1e10 if si else 1.0)
    pl.ux = rng.normal(0, 0.5, n) * pc.c
    pl.uy = rng.normal(0, 0.5, n) * pc.c
    pl.psi = rng.uniform(0.3, 2.0, n)
    pl.psi[::97] = 0.05           # a few QSA violators (gamma/psi > 35)
    pl.x_prev, pl.y_prev = pl.x.copy(), pl.y.copy()
    pl.ux_half, pl.uy_half, pl.psi_half = pl.ux.copy(), pl.uy.copy(), pl.psi.copy()
    pl.valid = np.ones(n, dtype=bool)
    pl.valid[::53] = False
    return geom, pc, F, pl, names


def _soa(pl):
    r = [pl.x, pl.y, pl.w, pl.ux, pl.uy, pl.psi, pl.x_prev, pl.y_prev, pl.ux_half, pl.uy_half, pl.psi_half]
    return [np.ascontiguousarray(a.copy()) for a in r]


def _grid(geom):
    g = geom.g
    return HcGrid(geom.nx + 2 * g, geom.ny + 2 * g, g, geom.pos_offset(0), geom.pos_offset(1),
                  1.0 / geom.dx, 1.0 / geom.dy)


def _planes(F, names):
    return np.ascontiguousarray(np.stack([F[('This', nm)] for nm in names]))


def _close(a, b, rtol=2e-13):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() <= rtol * scale


@pytest.mark.parametrize('si', [False, True])
@pytest.mark.parametrize('laser', [False, True])
@pytest.mark.parametrize('order', [0, 1, 2, 3])
def test_device_deposit_current_matches_oracle(hc, order, laser, si):
    geom, pc, F, pl, names = _setup(order, 2, laser, 10 + order, si)
    r = _soa(pl)
    valid = np.ascontiguousarray(pl.valid.astype(np.uint8))
    planes = _planes(F, names)
    ix = {nm: k for k, nm in enumerate(names)}
    c5 = (C.c_int * 6)(ix['jx'], ix['jy'], ix['jz'], ix['rho'], ix['chi'], ix['rhomjz'])
    invvol = 1.0 if not si else 1.0 / (geom.dx * geom.dy * geom.dz)
    norm = (pl.charge / pc.q_e) * (pc.m_e / pl.mass) * (pl.charge / pc.q_e) * (pc.m_e / pl.mass)
    par = DepositPar(1.0 / pc.c, pl.charge * invvol, pl.charge * pc.mu0 / pl.mass, 35.0, norm,
                     ix['aabs'] if laser else -1, pc.c)
    hg = _grid(geom)
    n_bad = hc.hc_deposit_current(order, C.c_long(pl.x.size), _ptrs(r), _dp(valid), _dp(planes),
                                  C.byref(hg), c5, C.byref(par))
    T = lambda nm: F[('This', nm)]
    want_bad = O.deposit_current(pl, F, geom, pc, not si, jx=T('jx'), jy=T('jy'), rho=T('rho'),
                                 chi=T('chi'), rhomjz=T('rhomjz'), aabs=T('aabs') if laser else None,
                                 jz=T('jz'))
    assert n_bad == want_bad and want_bad > 0
    assert np.array_equal(valid.astype(bool), pl.valid)
    for nm in ('jx', 'jy', 'jz', 'rho', 'chi', 'rhomjz'):
        assert _close(planes[ix[nm]], T(nm)), nm


@pytest.mark.parametrize('laser', [False, True])
@pytest.mark.parametrize('order,dtype', [(o, d) for o in range(4) for d in range(3) if (o, d) != (0, 0)])
def test_device_explicit_deposition_matches_oracle(hc, order, dtype, laser):
    for si in (False, True):
        geom, pc, F, pl, names = _setup(order, dtype, laser, 20 + order, si)
        r = _soa(pl)
        valid = np.ascontiguousarray(pl.valid.astype(np.uint8))
        planes = _planes(F, names)
        ix = {nm: k for k, nm in enumerate(names)}
        invvol = 1.0 if not si else 1.0 / (geom.dx * geom.dy * geom.dz)
        par = ExplicitPar(ix['Sy'], ix['Sx'], ix['Bz'], ix['Ez'], ix['ExmBy'], ix['EypBx'],
                          ix['aabs'] if laser else -1, pc.c, 1.0 / pc.c, pl.charge * invvol * pc.mu0,
                          pl.charge / pl.mass, (pc.m_e / pc.q_e) * (pc.m_e / pc.q_e))
        hg = _grid(geom)
        assert hc.hc_explicit_deposition(order, dtype, C.c_long(pl.x.size), _ptrs(r), _dp(valid),
                                         _dp(planes), C.byref(hg), C.byref(par)) == 0
        O.explicit_deposition(pl, F, geom, pc, not si, aabs=F[('This', 'aabs')] if laser else None)
        for nm in ('Sy', 'Sx'):
            assert _close(planes[ix[nm]], F[('This', nm)]), (nm, si)
            assert np.abs(F[('This', nm)]).max() > 0


@pytest.mark.parametrize('bc', ['Periodic', 'Reflecting', 'Absorbing'])
@pytest.mark.parametrize('laser', [False, True])
@pytest.mark.parametrize('order', [0, 1, 2, 3])
def test_device_advance_plasma_matches_oracle(hc, order, laser, bc):
    for si, nsubc, temp in ((False, 1, False), (True, 1, False), (False, 3, False), (False, 1, True)):
        geom, pc, F, pl, names = _setup(order, 2, laser, 30 + order, si)
        pl.n_subcycles = nsubc
        # fields weak enough for the particles to stay sane, strong enough to move them a few cells
        for nm in ('ExmBy', 'EypBx', 'Ez', 'Psi', 'Bx', 'By', 'Bz'):
            F[('This', nm)] *= 0.3
        pl.psi_half = np.random.default_rng(1).uniform(0.6, 2.0, pl.x.size)
        # shrink the particle box so that some particles leave it
        bc_lo = [geom.lo[0] + 2 * geom.dx, geom.lo[1] + 2 * geom.dy]
        bc_hi = [geom.hi[0] - 2 * geom.dx, geom.hi[1] - 2 * geom.dy]
        r = _soa(pl)
        valid = np.ascontiguousarray(pl.valid.astype(np.uint8))
        planes = _planes(F, names)
        ix = {nm: k for k, nm in enumerate(names)}
        norm = (pl.charge / pc.q_e) * (pc.m_e / pl.mass) * (pl.charge / pc.q_e) * (pc.m_e / pl.mass)
        par = PushPar(ix['Psi'], ix['Ez'], ix['Bx'], ix['By'], ix['Bz'], ix['aabs'] if laser else -1,
                      pc.c, pl.charge / (pl.mass * pc.c), geom.dz / nsubc, nsubc, int(temp),
                      {'Reflecting': 0, 'Periodic': 1, 'Absorbing': 2}[bc],
                      bc_lo[0], bc_lo[1], bc_hi[0], bc_hi[1], norm)
        hg = _grid(geom)
        assert hc.hc_advance_plasma(order, C.c_long(pl.x.size), _ptrs(r), _dp(valid), _dp(planes),
                                    C.byref(hg), C.byref(par)) == 0
        O.advance_plasma_particles(pl, F, geom, pc, bc, bc_lo, bc_hi, temp_slice=temp,
                                   aabs=F[('This', 'aabs')] if laser else None)
        assert np.array_equal(valid.astype(bool), pl.valid), (si, nsubc, temp)
        v = pl.valid
        want = [pl.x, pl.y, pl.w, pl.ux, pl.uy, pl.psi, pl.x_prev, pl.y_prev, pl.ux_half, pl.uy_half, pl.psi_half]
        for k, (got, w) in enumerate(zip(r, want)):
            assert _close(got[v], w[v], 1e-11), (k, si, nsubc, temp)
        assert np.array_equal(r[2], pl.w)          # weights: zeroed exactly where the oracle zeroes them


@needs_ref
@pytest.mark.parametrize('order', [0, 1, 2, 3])
def test_device_gather_matches_reference_header(hc, order):
    geom, pc, F, pl, names = _setup(order, 2, True, 40 + order, True)
    planes = _planes(F, names)
    ix = {nm: k for k, nm in enumerate(names)}
    comps = (C.c_int * 5)(ix['Psi'], ix['Ez'], ix['Bx'], ix['By'], ix['Bz'])
    hg = _grid(geom)
    n = pl.x.size
    out = np.zeros((6, n))
    x, y = np.ascontiguousarray(pl.x), np.ascontiguousarray(pl.y)
    assert hc.hc_gather(order, C.c_long(n), _dp(x), _dp(y), _dp(planes), C.byref(hg), comps, _dp(out)) == 0
    want = refhdr.ref_gather(order, x, y, planes, geom.g, [ix['Psi'], ix['Ez'], ix['Bx'], ix['By'], ix['Bz']],
                             1.0 / geom.dx, 1.0 / geom.dy, geom.pos_offset(0), geom.pos_offset(1))
    for k in range(6):
        assert _close(out[k], want[k], 1e-13), k
    ab = np.ascontiguousarray(F[('This', 'aabs')])
    out = np.zeros((4, n))
    assert hc.hc_laser_gather(order, C.c_long(n), _dp(x), _dp(y), _dp(ab), C.byref(hg), _dp(out)) == 0
    want = refhdr.ref_laser_gather(order, x, y, ab, geom.g, 1.0 / geom.dx, 1.0 / geom.dy,
                                   geom.pos_offset(0), geom.pos_offset(1))
    for k in range(4):
        assert _close(out[k], want[k], 1e-13), k


@pytest.mark.parametrize('order', [0, 1, 2, 3])
def test_device_beam_deposit_matches_oracle(hc, order):
    geom, pc, F, pl, names = _setup(order, 2, False, 50 + order, False)
    rng = np.random.default_rng(2)
    n = 2000
    bs = {'x': rng.uniform(geom.lo[0] + geom.dx, geom.hi[0] - geom.dx, n),
          'y': rng.uniform(geom.lo[1] + geom.dy, geom.hi[1] - geom.dy, n), 'z': np.zeros(n),
          'w': rng.uniform(0.5, 2, n), 'ux': rng.normal(0, 1, n), 'uy': rng.normal(0, 1, n),
          'uz': rng.normal(1000, 10, n), 'valid': rng.uniform(0, 1, n) > 0.1}
    beam = O.Beam('b', -1.0, 1.0, (1, 1, 1), 'flattop', 1.0, 0., 0., 1.)
    planes = np.zeros((3, geom.ny + 2 * geom.g, geom.nx + 2 * geom.g))
    b7 = [np.ascontiguousarray(bs[k]) for k in ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz')]
    valid = np.ascontiguousarray(bs['valid'].astype(np.uint8))
    hg = _grid(geom)
    assert hc.hc_beam_deposit(order, C.c_long(n), _ptrs(b7), _dp(valid), _dp(planes), C.byref(hg), 0, 1, 2,
                              C.c_double(1.0 / (pc.c * pc.c)), C.c_double(beam.charge)) == 0
    want = np.zeros_like(planes)
    O.beam_deposit(bs, beam, geom, pc, True, jxb=want[0], jyb=want[1], jzb=want[2])
    for k in range(3):
        assert _close(planes[k], want[k]), k


def test_only_order_3_deposits_depend_on_the_summation_order(hc):
    """Lattice particles (one per cell centre) deposited in two different particle orders: the
    weights of orders 0..2 are dyadic there (1; 1/2; 1/8, 3/4) so the sums are exact in any order;
    the order-3 weights (1/48, 23/48) are not, and the two rhomjz planes differ in the last bit.
    On the GPU the fp64 atomics arrive in no fixed order, which is why the neutral head slice of
    an order-3 run is 1e-16 noise instead of 0.0 and hpmg spends V-cycles on it
    (tests/test_gpu_zz_orders.py)."""
    rng = np.random.default_rng(1)
    for order in range(4):
        geom = O.Geometry(32, 32, 100, (-8., -8., -6.), (8., 8., 6.), order, 2)
        n, g = 32 * 32, geom.g
        ii, jj = np.meshgrid(np.arange(32), np.arange(32))
        x = geom.lo[0] + (ii.ravel() + 0.5) * geom.dx
        y = geom.lo[1] + (jj.ravel() + 0.5) * geom.dy
        res = []
        for perm in (np.arange(n), rng.permutation(n)):
            r = [np.ascontiguousarray(a) for a in (x[perm], y[perm], np.ones(n), np.zeros(n), np.zeros(n),
                                                   np.ones(n))] + [np.zeros(n) for _ in range(5)]
            valid = np.ones(n, dtype=np.uint8)
            planes = np.zeros((1, 32 + 2 * g, 32 + 2 * g))
            c5 = (C.c_int * 6)(-1, -1, -1, -1, -1, 0)
            par = DepositPar(1.0, -1.0, -1.0, 35.0, 1.0, -1, 1.0)
            hg = _grid(geom)
            hc.hc_deposit_current(order, C.c_long(n), _ptrs(r), _dp(valid), _dp(planes), C.byref(hg), c5,
                                  C.byref(par))
            res.append(planes[0, g:-g, g:-g].copy())
        diff = np.abs(res[0] - res[1]).max()
        assert (diff == 0.0) if order < 3 else (0.0 < diff < 1e-15), (order, diff)


class BxByRhsPar(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('c_jz', 'c_prev_jx', 'c_prev_jy', 'c_next_jx', 'c_next_jy')] + \
               [(k, C.c_double) for k in ('mu0', 'dx_inv_half', 'dy_inv_half', 'dz_inv_half')]


def test_device_open_boundary_and_bxby_rhs_match_oracle(hc):
    """pc_fields.cuh on the host: the multipole moments + edge values of boundary.field = Open and
    the right-hand sides of SolvePoissonBxBy against the oracle (which the reference's
    beam_in_vacuum_open_boundary golden pins)"""
    rng = np.random.default_rng(12)
    geom = O.Geometry(48, 36, 10, (-4.0, -3.0, -2.0), (5.0, 3.5, 2.0), 2, 2)
    nx, ny = geom.nx, geom.ny
    for monopole in (1, 0):
        rhs = rng.normal(0, 1, (ny, nx))
        want = O.open_boundary_rhs(rhs, geom, bool(monopole))
        got = np.ascontiguousarray(rhs.copy())
        assert hc.hc_open_boundary(_dp(got), nx, ny, C.c_double(geom.dx), C.c_double(geom.dy),
                                   C.c_double(geom.lo[0]), C.c_double(geom.hi[0]), C.c_double(geom.lo[1]),
                                   C.c_double(geom.hi[1]), monopole) == 0
        assert np.abs(want - rhs).max() > 0                      # the edges did change
        assert np.abs(got - want).max() <= 1e-11 * np.abs(want - rhs).max()
        assert np.array_equal(got[1:-1, 1:-1], rhs[1:-1, 1:-1])  # ... and only the edges
    g = geom.g
    names = ['jz', 'pjx', 'pjy', 'njx', 'njy']
    planes = np.ascontiguousarray(rng.normal(0, 1, (5, ny + 2 * g, nx + 2 * g)))
    par = BxByRhsPar(0, 1, 2, 3, 4, 1.3, 0.5 / geom.dx, 0.5 / geom.dy, 0.5 / geom.dz)
    stage = np.zeros((2, ny, nx))
    hg = _grid(geom)
    hc.hc_bxby_rhs(_dp(planes), C.byref(hg), C.byref(par), nx, ny, _dp(stage))
    v = (slice(g, -g), slice(g, -g))
    dz_jy = (planes[2][v] - planes[4][v]) * (0.5 / geom.dz)
    dz_jx = (planes[1][v] - planes[3][v]) * (0.5 / geom.dz)
    want_bx = -1.3 * O._ddy(planes[0], geom.dy, g) + 1.3 * dz_jy
    want_by = 1.3 * O._ddx(planes[0], geom.dx, g) + (-1.3) * dz_jx
    assert np.abs(stage[0] - want_bx).max() <= 1e-13 * np.abs(want_bx).max()
    assert np.abs(stage[1] - want_by).max() <= 1e-13 * np.abs(want_by).max()


class LaserAdvPar(C.Structure):
    _fields_ = [('nx', C.c_int), ('ny', C.c_int), ('step0', C.c_int)] + \
               [(k, C.c_double) for k in ('dx', 'dy', 'dz', 'c', 'dt', 'k0', 'dkx', 'dky')]


@pytest.mark.parametrize('step0', [1, 0])
@pytest.mark.parametrize('nx,ny', [(32, 24), (17, 21)])
def test_device_laser_advance_matches_oracle(hc, step0, nx, ny):
    """laser_advance.cuh on the host: right-hand side, on-axis phase and spectral division of
    AdvanceSliceFFT around NumPy's FFT reproduce the oracle's laser_advance_fft (which the
    reference's laser_evolution golden pins); InterpolateChi and UpdateLaserAabs from a stored slice"""
    rng = np.random.default_rng(21 + nx)
    geom = O.Geometry(nx, ny, 10, (-6e-5, -5e-5, -8e-5), (6e-5, 5e-5, 6e-5), 0, 2)
    pc = O.PhysConst.make(False)
    lam, dt = 0.8e-6, 70e-6 / pc.c
    L = O.LaserSlices(ny, nx)
    x = (np.arange(nx) - nx / 2 + 0.5)[None, :] / nx
    y = (np.arange(ny) - ny / 2 + 0.5)[:, None] / ny
    env = np.exp(-20 * (x * x + y * y))
    names = ('nm1j00', 'nm1jp1', 'nm1jp2', 'n00j00', 'n00jp1', 'n00jp2', 'np1jp1', 'np1jp2')
    for k, nm in enumerate(names):
        setattr(L, nm, env * np.exp(1j * (0.3 * k + 2.0 * x)) * (1 + 0.05 * rng.normal(size=(ny, nx))))
    chi = rng.uniform(0, 1e9, (ny, nx))
    O.laser_advance_fft(L, chi, geom, pc, lam, dt, 0 if step0 else 3, True)
    want = L.np1j00
    # the same through the device functions
    par = LaserAdvPar(nx, ny, step0, geom.dx, geom.dy, geom.dz, pc.c, dt, 2 * np.pi / lam,
                      2 * np.pi / (geom.hi[0] - geom.lo[0]), 2 * np.pi / (geom.hi[1] - geom.lo[1]))
    planes = [np.ascontiguousarray(getattr(L, nm)) for nm in names]
    imid, jmid = (nx + 1) // 2, (ny + 1) // 2
    kx = [imid - 1, imid] if nx % 2 == 0 else [imid]
    ky = [jmid - 1, jmid] if ny % 2 == 0 else [jmid]
    h3 = np.array([[h.real, h.imag] for h in (L.n00j00[np.ix_(ky, kx)].sum(), L.n00jp1[np.ix_(ky, kx)].sum(),
                                              L.n00jp2[np.ix_(ky, kx)].sum())]).ravel()
    rhs = np.zeros((ny, nx), dtype=complex)
    phase = np.zeros(5)
    chi_c = np.ascontiguousarray(chi)
    hc.hc_laser_rhs(_ptrs(planes), _dp(chi_c), C.byref(par), _dp(h3), 1, _dp(rhs), _dp(phase))
    rhs2, phase2 = np.zeros((ny, nx), dtype=complex), np.zeros(5)      # on-axis sums by the device function
    hc.hc_laser_rhs(_ptrs(planes), _dp(chi_c), C.byref(par), None, 1, _dp(rhs2), _dp(phase2))
    assert np.allclose(phase, phase2, rtol=1e-12, atol=0) and np.abs(rhs - rhs2).max() <= 1e-12 * np.abs(rhs).max()
    hc.hc_laser_diag_xz_sum.restype = C.c_double
    env0 = planes[3]
    line = 0.5 * (env0[ny // 2 - 1] + env0[ny // 2]) if ny % 2 == 0 else env0[ny // 2]
    assert hc.hc_laser_diag_xz_sum(_dp(env0), nx, ny) == pytest.approx(np.abs(line).sum(), rel=1e-14)
    rhs_f = np.ascontiguousarray(np.fft.fft2(rhs))
    hc.hc_laser_spectral(_dp(rhs_f), C.byref(par), _dp(phase))
    got = np.fft.ifft2(rhs_f) * (nx * ny)                 # cuFFT's inverse is unnormalised
    assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max()
    assert np.abs(want).max() > 0


def test_device_laser_chi_and_aabs_match_oracle(hc):
    rng = np.random.default_rng(30)
    for order in (0, 1, 2):
        geom = O.Geometry(20, 26, 8, (-3e-5, -2e-5, 0.0), (5e-5, 7e-5, 1e-5), 2, 2)
        g = geom.g
        chi_field = rng.uniform(0, 1, (geom.ny + 2 * g, geom.nx + 2 * g))
        chi_init = rng.uniform(0, 1, (geom.ny, geom.nx))
        env = rng.normal(size=(geom.ny, geom.nx)) + 1j * rng.normal(size=(geom.ny, geom.nx))
        want_chi = O.laser_interpolate_chi(chi_field, chi_init, geom, order)
        want_aabs = np.zeros_like(chi_field)
        O.update_laser_aabs(env, want_aabs, geom, order)
        planes = np.ascontiguousarray(chi_field[None])
        hg = _grid(geom)
        chi_out = np.zeros((geom.ny, geom.nx))
        aabs_out = np.zeros_like(chi_field)
        ci, ev = np.ascontiguousarray(chi_init), np.ascontiguousarray(env)
        hc.hc_laser_chi_aabs(_dp(planes), C.byref(hg), 0, _dp(ci), _dp(ev), geom.nx, geom.ny,
                             C.c_double(geom.dx), C.c_double(geom.dy), order, _dp(chi_out), _dp(aabs_out))
        assert np.abs(chi_out - want_chi).max() <= 1e-14
        assert np.abs(aabs_out - want_aabs).max() <= 1e-13 * np.abs(want_aabs).max()
