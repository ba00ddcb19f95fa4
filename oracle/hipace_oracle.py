"""CPU oracle: a NumPy/SciPy fp64 restatement of the HiPACE++ per-zeta-slice quasi-static PIC loop.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product path
(hipace_b200/ + libhpb200.so) never calls into this file.

Parity status: PINNED.  tests/test_oracle_golden.py checks the checksums this oracle produces
against FOURTEEN of the reference's own golden files (tests/checksum/benchmarks_json/, rtol 1e-9;
copies of the numbers are committed under tests/golden/ with the script that extracted them) --
every golden of the reference that does not depend on AMReX's random number streams:
  linear_wake.{normalized,SI}.1Rank, gaussian_linear_wake.{normalized,SI}.1Rank,
  blowout_wake_explicit.2Rank, beam_evolution.1Rank (20 time steps), beam_in_vacuum.{normalized,SI}.1Rank
  (deposition order 0), grid_current.1Rank, adaptive_time_step.1Rank (hipace.dt = adaptive),
  beam_in_vacuum_open_boundary.normalized.1Rank (predictor-corrector solver, open field boundaries),
  laser_blowout_wake_explicit.{1Rank,SI.1Rank} (laser at step 0), laser_evolution.SI.2Rank (envelope advance).
tests/test_oracle_refheaders.py additionally holds the shape factors (every order and derivative
type), the field / laser gathers, the momentum derivative (plain and dual, bit-identical) and the
open-boundary multipoles to the reference's OWN headers, compiled in place by oracle/ref_headers.cpp.

Each function cites the reference file:line (relative to /root/reference/src) it restates.
Scope (SURVEY.md section 8 and its "next" rows): level 0; explicit (hpmg) and predictor-corrector
Bx/By solvers; depos_order_xy 0..3, depos_derivative_type 0..2; Dirichlet and Open field
boundaries; gaussian lasers incl. the envelope advance (fft and multigrid solvers); fixed and
adaptive time step; in-situ diagnostics; no ionization, no MR; fixed_ppc beams, u_std = 0 plasma (no
RNG anywhere).  Not pinned by any golden: the hpmg type 2 laser solver (held to the fft solver), the
plasma half of the predictor-corrector loop (held to the explicit solver).

Array convention: every slice component is a 2-D array a[j + g, i + g] (x fastest) over the
grown box [-g, n-1+g]^2, g = (depos_order_xy + 1) / 2 + 1 guard cells (fields/Fields.cpp:63-64).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import numpy as np

try:  # SciPy is only needed for the Poisson solve
    from scipy.fft import dstn as _dstn
except Exception:  # pragma: no cover
    _dstn = None

G = 2  # guard cells of the DEFAULT depos_order_xy = 2 (Geometry.g holds the deck's; this constant
       # is what the order-2-only C port and the CUDA parity tests index with)


# --------------------------------------------------------------------------------------------
# Input deck (subset of AMReX ParmParse + HiPACE++ parser, utils/Parser.H:316-395)
# --------------------------------------------------------------------------------------------

_CONST_SI = dict(clight=299792458.0, epsilon0=8.8541878128e-12, mu0=1.25663706212e-06,
                 q_e=1.602176634e-19, m_e=9.1093837015e-31, m_p=1.67262192369e-27,
                 hbar=1.054571817e-34, r_e=2.817940326204929e-15, pi=math.pi)


def parse_deck(text: str, overrides: dict | None = None) -> dict:
    """'prefix.key = v1 v2 ...' lines, '#' comments; later entries win (like CLI overrides)."""
    deck: dict[str, list[str]] = {}
    for raw in text.splitlines():
        line = raw.split('#', 1)[0].strip()
        if not line or '=' not in line:
            continue
        key, val = line.split('=', 1)
        toks = re.findall(r'"[^"]*"|\S+', val.strip())
        deck[key.strip()] = [t.strip('"') for t in toks]
    for k, v in (overrides or {}).items():
        deck[k] = [str(t) for t in (v if isinstance(v, (list, tuple)) else str(v).split())]
    return deck


def _constants(deck: dict) -> dict:
    """my_constants.* may refer to each other in any order (utils/Parser.H:134-170): resolve by
    repeated passes until nothing new can be evaluated"""
    env = dict(_CONST_SI)
    env.update({f: getattr(math, f) for f in ('sqrt', 'exp', 'sin', 'cos', 'log', 'tanh')})
    # (an expression may contain blanks: the deck reader split it into tokens)
    todo = {k.split('.', 1)[1]: ' '.join(v) for k, v in deck.items() if k.startswith('my_constants.')}
    while todo:
        done = []
        for name, expr in todo.items():
            try:
                env[name] = float(eval(expr.replace('^', '**'), {'__builtins__': {}}, env))
                done.append(name)
            except NameError:
                pass
        if not done:
            raise ValueError('unresolvable my_constants: ' + ', '.join(todo))
        for name in done:
            del todo[name]
    return env


def _eval(expr: str, deck: dict, extra: dict | None = None) -> float:
    """Math-parser stand-in: numbers, + - * / ^, my_constants.*, built-in constants."""
    env = _constants(deck)
    env.update(extra or {})
    return float(eval(expr.replace('^', '**'), {'__builtins__': {}}, env))


def _get(deck, key, default=None, n=None, typ=float, alt=None):
    v = deck.get(key)
    if v is None and alt is not None:
        v = deck.get(alt)
    if v is None:
        return default
    if typ is str:
        return v[0] if n is None else v
    if n is None:       # a scalar: the whole right-hand side is ONE expression, blanks included
        return typ(_eval(' '.join(v), deck))                # (utils/Parser.H: getWithParser joins the tokens)
    return [typ(_eval(t, deck)) for t in v]


@dataclass
class PhysConst:
    """utils/Constants.H:54-81"""
    c: float
    ep0: float
    mu0: float
    q_e: float
    m_e: float
    m_p: float

    @staticmethod
    def make(normalized: bool) -> 'PhysConst':
        if normalized:
            return PhysConst(1., 1., 1., 1., 1., 1836.15267343)
        s = _CONST_SI
        return PhysConst(s['clight'], s['epsilon0'], s['mu0'], s['q_e'], s['m_e'], s['m_p'])


# --------------------------------------------------------------------------------------------
# Shape factors (particles/particles_utils/ShapeFactors.H)
# --------------------------------------------------------------------------------------------

def shape_order2(xmid):
    """compute_single_shape_factor<false,2>, ShapeFactors.H:165-174.
    Returns (S[3,P], leftmost cell[P])."""
    xfloor = np.floor(xmid + 0.5)
    xint = xmid - xfloor
    s = np.stack([0.5 * (0.5 - xint) * (0.5 - xint),
                  0.75 - xint * xint,
                  0.5 * (0.5 + xint) * (0.5 + xint)])
    return s, xfloor.astype(np.int64) - 1


def dshape_centered_order2(xmid):
    """single_derivative_shape_factor<2,2>, ShapeFactors.H:405-430.
    Returns (S[5,P], dS[5,P] (already '-sdx'), leftmost cell[P])."""
    xm = xmid + 0.5
    xfloor = np.floor(xm)
    xint = xm - xfloor
    x2 = xint * xint
    z = np.zeros_like(xint)
    s = np.stack([z,
                  0.5 * x2 - xint + 0.5,
                  -x2 + xint + 0.5,
                  0.5 * x2,
                  z])
    sdx = np.stack([-0.25 * x2 + 0.5 * xint - 0.25,
                    0.5 * x2 - 0.5 * xint - 0.25,
                    0.25 - 0.5 * xint,
                    -0.5 * x2 + 0.5 * xint + 0.25,
                    0.25 * x2])
    return s, -sdx, xfloor.astype(np.int64) - 2


def dshape_nodal_order2(xmid):
    """single_derivative_shape_factor<1,2>, ShapeFactors.H:305-329.
    Returns (S[4,P], dS[4,P] (already '-sdx'), leftmost cell[P])."""
    xfloor = np.floor(xmid)
    xint = xmid - xfloor
    x2 = xint * xint
    lo = xint < 0.5
    z = np.zeros_like(xint)
    s = np.stack([np.where(lo, 0.5 * x2 - 0.5 * xint + 0.125, z),
                  np.where(lo, 0.75 - x2, 0.5 * x2 - 1.5 * xint + 1.125),
                  np.where(lo, 0.5 * x2 + 0.5 * xint + 0.125, -x2 + 2 * xint - 0.25),
                  np.where(lo, z, 0.5 * x2 - 0.5 * xint + 0.125)])
    sdx = np.stack([-0.5 * x2 + xint - 0.5,
                    1.5 * x2 - 2 * xint,
                    -1.5 * x2 + xint + 0.5,
                    0.5 * x2])
    return s, -sdx, xfloor.astype(np.int64) - 1


def _bspline_pieces(nmax=3):
    """Polynomial pieces of the uniform B-spline B_n on the knots 0..n+1 (Cox-de Boor):
    B_n(i + f) = P[n][i](f), 0 <= f < 1 -- the closed form behind every table of ShapeFactors.H,
    S_n(x) = B_n(x + (n+1)/2).  Coefficients low order first."""
    from numpy.polynomial import polynomial as npoly
    P = {0: [np.array([1.0])]}
    for n in range(1, nmax + 1):
        P[n] = []
        for i in range(n + 1):
            acc = np.zeros(n + 1)
            if i <= n - 1:
                acc = npoly.polyadd(acc, npoly.polymul([i / n, 1.0 / n], P[n - 1][i]))
            if i >= 1:
                acc = npoly.polyadd(acc, npoly.polymul([(n + 1 - i) / n, -1.0 / n], P[n - 1][i - 1]))
            P[n].append(acc)
    return P


_BP = _bspline_pieces()


def _piece(n, i, f, deriv=False):
    """B_n (or its derivative) on its piece i at local coordinate f; 0 outside the support"""
    if i < 0 or i > n:
        return np.zeros_like(f)
    c = _BP[n][i]
    if deriv:
        c = c[1:] * np.arange(1, len(c))
        if len(c) == 0:
            return np.zeros_like(f)
    out = np.full_like(f, c[-1])
    for a in c[-2::-1]:
        out = out * f + a
    return out


def _bspline_at(n, halves, xint, deriv=False):
    """B_n at t = xint + halves/2, 0 <= xint < 1: a whole number of cells picks the piece outright;
    a half-cell offset needs the one test the reference's tables have, xint < 0.5"""
    if halves % 2 == 0:
        return _piece(n, halves // 2, xint, deriv)
    return np.where(xint < 0.5, _piece(n, (halves - 1) // 2, xint + 0.5, deriv),
                    _piece(n, (halves + 1) // 2, xint - 0.5, deriv))


def _cell_and_xint(m, xmid):
    """leftmost cell of an order-m stencil and the in-cell coordinate it is counted from:
    even m hang on the nearest cell (floor(xmid + 1/2)), odd m on floor(xmid)"""
    xm = xmid + 0.5 if m % 2 == 0 else xmid
    xfloor = np.floor(xm)
    return xfloor.astype(np.int64) - (m // 2 if m % 2 == 0 else (m - 1) // 2), xm - xfloor


def shape(order, xmid):
    """compute_shape_factor<order> / compute_single_shape_factor<.,order>, ShapeFactors.H:40-117,
    122-195: (weights[order+1, P], leftmost cell[P]).  Order 2 keeps the literal polynomials."""
    if order == 2:
        return shape_order2(xmid)
    j0, xint = _cell_and_xint(order, xmid)
    return np.stack([_bspline_at(order, 2 * (order - k), xint) for k in range(order + 1)]), j0


def dshape(dtype, order, xmid):
    """single_derivative_shape_factor<dtype, order>, ShapeFactors.H:211-466:
    (S[m, P], dS[m, P] (the '-sdx' the reference returns), leftmost cell[P]), m = order+dtype+1.
    The tables of the reference are, per cell at distance x = xmid - cell,
      weight S_order(x) and
      type 0 (analytic):  -S_order'(x)
      type 1 (nodal):      S_order(x - 1/2) - S_order(x + 1/2)     (staggered difference)
      type 2 (centred):   (S_order(x - 1) - S_order(x + 1)) / 2    (central difference)
    with the stencil of order+1 for type 1 and the order's own stencil grown by one cell on both
    sides for type 2.  Pinned to the reference's header itself in tests/test_oracle_refheaders.py."""
    if order == 2 and dtype == 2:
        return dshape_centered_order2(xmid)
    if order == 2 and dtype == 1:
        return dshape_nodal_order2(xmid)
    m = order + 1 if dtype == 1 else order
    e = 1 if dtype == 2 else 0
    j0, xint = _cell_and_xint(m, xmid)
    j0 = j0 - e
    s, ds = [], []
    for k in range(order + dtype + 1):
        h = m + order + 2 * e - 2 * k          # t = x + (order+1)/2 = xint + h/2
        s.append(_bspline_at(order, h, xint))
        if dtype == 0:
            ds.append(-_bspline_at(order, h, xint, deriv=True))
        elif dtype == 1:
            ds.append(_bspline_at(order, h - 1, xint) - _bspline_at(order, h + 1, xint))
        else:
            ds.append(0.5 * (_bspline_at(order, h - 2, xint) - _bspline_at(order, h + 2, xint)))
    return np.stack(s), np.stack(ds), j0


def n_guards(order):
    """fields/Fields.cpp:63-64"""
    return (order + 1) // 2 + 1


# --------------------------------------------------------------------------------------------
# Geometry and field storage
# --------------------------------------------------------------------------------------------

@dataclass
class Geometry:
    nx: int
    ny: int
    nz: int
    lo: tuple
    hi: tuple
    order: int = 2      # hipace.depos_order_xy
    dtype: int = 2      # hipace.depos_derivative_type

    @property
    def g(self):
        return n_guards(self.order)

    @property
    def dx(self):
        return (self.hi[0] - self.lo[0]) / self.nx

    @property
    def dy(self):
        return (self.hi[1] - self.lo[1]) / self.ny

    @property
    def dz(self):
        return (self.hi[2] - self.lo[2]) / self.nz

    def pos_offset(self, d):
        """GetPosOffset, fields/Fields.H:71-77, with the grown box [-g, n-1+g] (or the valid box:
        both give lo + dx/2 up to round-off; we follow the formula literally)."""
        n = (self.nx, self.ny, self.nz)[d]
        dd = (self.dx, self.dy, self.dz)[d]
        g = self.g if d < 2 else 0
        return 0.5 * (self.lo[d] + self.hi[d] - dd * ((-g) + (n - 1 + g)))


# component names of the explicit solver, allocation order of fields/Fields.cpp:70-122
EXPLICIT_COMPS = (
    ('Next', ('jx_beam', 'jy_beam')),
    ('This', ('chi', 'Sy', 'Sx', 'ExmBy', 'EypBx', 'Ez', 'Bx', 'By', 'Bz', 'Psi',
              'jx_beam', 'jy_beam', 'jz_beam', 'jx', 'jy', 'rhomjz')),
    ('Previous', ('jx_beam', 'jy_beam')),
    ('RhomJzIons', ('rhomjz',)),
)


PC_COMPS = (        # predictor-corrector solver, fields/Fields.cpp:124-163
    ('Next', ('jx', 'jy')),
    ('This', ('ExmBy', 'EypBx', 'Ez', 'Bx', 'By', 'Bz', 'Psi', 'jx', 'jy', 'jz', 'rhomjz')),
    ('Previous', ('Bx', 'By', 'jx', 'jy')),
    ('RhomJzIons', ('rhomjz',)),
    ('PCIter', ('Bx', 'By')),
    ('PCPrevIter', ('Bx', 'By')),
)


def component_map(deposit_rho=False, neutral_background=True, use_laser=False, explicit=True):
    comps = {}
    n = 0
    if not explicit:
        assert not use_laser, 'oracle scope: laser only with the explicit solver'
        for sl, names in PC_COMPS:
            if sl == 'RhomJzIons' and not neutral_background:
                continue
            for nm in names:
                comps[(sl, nm)] = n
                n += 1
            if sl == 'This' and deposit_rho:
                comps[('This', 'rho')] = n
                n += 1
        return comps, n
    for sl, names in EXPLICIT_COMPS:
        if sl == 'RhomJzIons' and not neutral_background:
            continue
        for nm in names:
            comps[(sl, nm)] = n
            n += 1
        if sl == 'This' and use_laser:                   # fields/Fields.cpp:98-101
            comps[('This', 'aabs')] = n
            n += 1
        if sl == 'This' and deposit_rho:
            comps[('This', 'rho')] = n
            n += 1
    return comps, n


# --------------------------------------------------------------------------------------------
# Species
# --------------------------------------------------------------------------------------------

@dataclass
class Plasma:
    name: str
    charge: float
    mass: float
    ppc: tuple
    density: object            # callable(x, y, z) -> array
    neutralize_background: bool = True
    max_qsa_weighting_factor: float = 35.
    n_subcycles: int = 1
    radius: float = math.inf
    hollow_core_radius: float = 0.
    min_density: float = 0.
    u_mean: tuple = (0., 0., 0.)
    # SoA, PlasmaIdx order particles/plasma/PlasmaParticleContainer.H:21-46
    x: np.ndarray = None
    y: np.ndarray = None
    w: np.ndarray = None
    ux: np.ndarray = None
    uy: np.ndarray = None
    psi: np.ndarray = None
    x_prev: np.ndarray = None
    y_prev: np.ndarray = None
    ux_half: np.ndarray = None
    uy_half: np.ndarray = None
    psi_half: np.ndarray = None
    valid: np.ndarray = None     # id sign (valid/invalid)


@dataclass
class Beam:
    name: str
    charge: float
    mass: float
    ppc: tuple
    profile: str
    density: float
    zmin: float
    zmax: float
    radius: float
    position_mean: tuple = (0., 0., 0.)
    position_std: tuple = (0., 0., 0.)
    u_mean: tuple = (0., 0., 0.)
    min_density: float = 0.
    n_subcycles: int = 10
    do_z_push: bool = True
    external_fields: object = None   # callable(x, y, z, t) -> (Ex, Ey, Ez, Bx, By, Bz) or None
    slices: dict = field(default_factory=dict)   # islice -> dict of arrays
    next_id: int = 1


# --------------------------------------------------------------------------------------------
# Kernels
# --------------------------------------------------------------------------------------------

# ---- laser envelope: analytic envelope at step 0, |a|^2 on the field grid, the
# ponderomotive terms of the particle kernels, and the envelope advance (fft and multigrid solvers)

@dataclass
class Laser:
    name: str
    a0: float = 0.0
    w0: float = 0.0
    cep: float = 0.0
    propagation_angle_yz: float = 0.0
    pft_yz: float = math.pi / 2.0
    L0: float = 0.0
    focal_distance: float = 0.0
    position_mean: tuple = (0.0, 0.0, 0.0)


def shape_order_n(xmid, order):
    """compute_shape_factor<order> for order 0..2 (particles_utils/ShapeFactors.H:40-117):
    weights and leftmost cell"""
    if order == 0:
        return [np.ones_like(xmid)], np.floor(xmid + 0.5).astype(np.int64)
    if order == 1:
        j = np.floor(xmid)
        xint = xmid - j
        return [1.0 - xint, xint], j.astype(np.int64)
    if order == 2:
        return shape_order2(xmid)
    raise NotImplementedError('interp order 3')


def laser_envelope_slice(lasers, lambda0, geom: 'Geometry', islice: int):
    """MultiLaser::InitLaserSlice, gaussian branch, laser/MultiLaser.cpp:881-917 (laser grid =
    field grid, the default of MakeLaserGeometry :58-118).  Complex array a[j, i], valid box."""
    k0 = 2.0 * math.pi / lambda0
    x = (np.arange(geom.nx) * geom.dx + geom.pos_offset(0))[None, :]
    y = (np.arange(geom.ny) * geom.dy + geom.pos_offset(1))[:, None]
    z = islice * geom.dz + geom.pos_offset(2)
    env = np.zeros((geom.ny, geom.nx), dtype=complex)
    for L in lasers:
        xs, ys, zs = x - L.position_mean[0], y - L.position_mean[1], z - L.position_mean[2]
        ang = L.propagation_angle_yz + (L.pft_yz - math.pi / 2.0)
        yp = math.cos(ang) * ys - math.sin(ang) * zs
        zp = math.sin(ang) * ys + math.cos(ang) * zs
        diffract = 1.0 + 1j * (zp - L.focal_distance + L.position_mean[2] * math.cos(L.propagation_angle_yz)) \
            * 2.0 / (k0 * L.w0 * L.w0)
        inv_waist2 = 1.0 / (L.w0 * L.w0 * diffract)
        prefactor = L.a0 / diffract
        stc = prefactor * np.exp(-(zp * zp / (L.L0 * L.L0)))
        arg = -(xs * xs + yp * yp) * inv_waist2
        env = env + stc * np.exp(arg) * np.exp(1j * yp * k0 * L.propagation_angle_yz + L.cep)
    return env


def update_laser_aabs(env, aabs, geom: 'Geometry', interp_order: int):
    """MultiLaser::UpdateLaserAabs, laser/MultiLaser.cpp:214-291: |a|^2 of the current envelope
    interpolated (order interp_order) from the laser grid to the grown field slice"""
    a2 = env.real * env.real + env.imag * env.imag                   # abssq
    G = geom.g
    ii = np.arange(-G, geom.nx + G)
    jj = np.arange(-G, geom.ny + G)
    # same geometry for both grids: xmid = ((i dx + off) - off) / dx
    xmid = ((ii * geom.dx + geom.pos_offset(0)) - geom.pos_offset(0)) * (1.0 / geom.dx)
    ymid = ((jj * geom.dy + geom.pos_offset(1)) - geom.pos_offset(1)) * (1.0 / geom.dy)
    sx, i0 = shape_order_n(xmid, interp_order)
    sy, j0 = shape_order_n(ymid, interp_order)
    out = np.zeros((geom.ny + 2 * G, geom.nx + 2 * G))
    for iy in range(interp_order + 1):
        for ix in range(interp_order + 1):
            cx, cy = i0 + ix, j0 + iy
            okx = (cx >= 0) & (cx <= geom.nx - 1)
            oky = (cy >= 0) & (cy <= geom.ny - 1)
            val = a2[np.clip(cy, 0, geom.ny - 1)[:, None], np.clip(cx, 0, geom.nx - 1)[None, :]]
            w = (sy[iy] * oky)[:, None] * (sx[ix] * okx)[None, :]
            out += w * val
    aabs[...] = out


class LaserSlices:
    """The nine complex work slices of MultiLaser (laser/MultiLaser.H:24-48, WhichLaserSlice):
    time levels n-1, n, n+1 at slices j, j+1, j+2, on the valid laser box."""

    def __init__(self, ny, nx):
        z = lambda: np.zeros((ny, nx), dtype=complex)
        self.nm1j00, self.nm1jp1, self.nm1jp2 = z(), z(), z()
        self.n00j00, self.n00jp1, self.n00jp2 = z(), z(), z()
        self.np1j00, self.np1jp1, self.np1jp2 = z(), z(), z()

    def shift(self):
        """MultiLaser::ShiftLaserSlices, laser/MultiLaser.cpp:180-212 (the j00 slots are then
        refilled by MultiBuffer::get_data / InitSliceEnvelope)"""
        self.nm1jp2, self.nm1jp1 = self.nm1jp1, self.nm1j00
        self.n00jp2, self.n00jp1 = self.n00jp1, self.n00j00
        self.np1jp2, self.np1jp1 = self.np1jp1, self.np1j00


def laser_interpolate_chi(chi_field, chi_initial, geom: 'Geometry', interp_order: int):
    """MultiLaser::InterpolateChi, laser/MultiLaser.cpp:334-407 for coinciding grids: chi of the
    field slice inside the field box shrunk by 2 guard widths, the initial chi outside"""
    G = geom.g
    ii, jj = np.arange(geom.nx), np.arange(geom.ny)
    xmid = ((ii * geom.dx + geom.pos_offset(0)) - geom.pos_offset(0)) * (1.0 / geom.dx)
    ymid = ((jj * geom.dy + geom.pos_offset(1)) - geom.pos_offset(1)) * (1.0 / geom.dy)
    sx, i0 = shape_order_n(xmid, interp_order)
    sy, j0 = shape_order_n(ymid, interp_order)
    chi = np.zeros((geom.ny, geom.nx))
    for iy in range(interp_order + 1):
        for ix in range(interp_order + 1):
            chi += (sy[iy][:, None] * sx[ix][None, :]) * chi_field[(j0 + iy + G)[:, None], (i0 + ix + G)[None, :]]
    # field box [-G, n-1+G] grown by -2G -> [G, n-1-G]; laser index range where it ends (:368-371)
    x_lo, x_hi, y_lo, y_hi = G, geom.nx - 1 - G, G, geom.ny - 1 - G
    inside = ((jj >= y_lo) & (jj <= y_hi))[:, None] & ((ii >= x_lo) & (ii <= x_hi))[None, :]
    return np.where(inside, chi, chi_initial)


def laser_advance_fft(L: 'LaserSlices', chi, geom: 'Geometry', pc: 'PhysConst', lambda0, dt, step,
                      use_phase=True):
    """MultiLaser::AdvanceSliceFFT, laser/MultiLaser.cpp:609-801: A^{n+1}_j from the envelope
    equation (Benedetti et al. 2017 discretisation), solved with a 2-D complex FFT"""
    dx, dy, dz, c = geom.dx, geom.dy, geom.dz, pc.c
    k0 = 2.0 * math.pi / lambda0
    Ny, Nx = L.n00j00.shape
    imid, jmid = (Nx + 1) // 2, (Ny + 1) // 2
    tj00 = tjp1 = tjp2 = 0.0
    if use_phase:                                                  # on-axis phase, :651-685
        keep_x = [imid - 1, imid] if Nx % 2 == 0 else [imid]
        keep_y = [jmid - 1, jmid] if Ny % 2 == 0 else [jmid]
        ax = lambda a: a[np.ix_(keep_y, keep_x)].sum()
        h0, h1, h2 = ax(L.n00j00), ax(L.n00jp1), ax(L.n00jp2)
        tj00, tjp1, tjp2 = (math.atan2(h.imag, h.real) for h in (h0, h1, h2))
    dt1, dt2 = tj00 - tjp1, tjp1 - tjp2
    if dt1 < -1.5 * math.pi: dt1 += 2.0 * math.pi
    if dt1 > 1.5 * math.pi: dt1 -= 2.0 * math.pi
    if dt2 < -1.5 * math.pi: dt2 += 2.0 * math.pi
    if dt2 > 1.5 * math.pi: dt2 -= 2.0 * math.pi
    exp1 = np.exp(1j * (tj00 - tjp1))
    exp2 = np.exp(1j * (tj00 - tjp2))
    djn = (-3.0 * dt1 + dt2) / (2.0 * dz)

    def lap(a):                                                    # 0 on the edge cells, :703-722
        out = np.zeros_like(a)
        out[1:-1, 1:-1] = (a[1:-1, 2:] + a[1:-1, :-2] - 2.0 * a[1:-1, 1:-1]) / (dx * dx) \
            + (a[2:, 1:-1] + a[:-2, 1:-1] - 2.0 * a[1:-1, 1:-1]) / (dy * dy)
        return out
    an00j00 = L.n00j00
    if step == 0:
        rhs = (8.0 / (c * dt * dz) * (-L.np1jp1 + L.n00jp1) * exp1
               + 2.0 / (c * dt * dz) * (L.np1jp2 - L.n00jp2) * exp2
               + 2.0 * chi * an00j00
               - lap(L.n00j00)
               + (-6.0 / (c * dt * dz) + 4.0 * 1j * djn / (c * dt) + 1j * 4.0 * k0 / (c * dt)) * an00j00)
        acoeff = 6.0 / (c * dt * dz) - 1j * 4.0 * (k0 + djn) / (c * dt)
    else:
        rhs = (4.0 / (c * dt * dz) * (-L.np1jp1 + L.nm1jp1) * exp1
               + 1.0 / (c * dt * dz) * (L.np1jp2 - L.nm1jp2) * exp2
               - 4.0 / (c * c * dt * dt) * an00j00
               + 2.0 * chi * an00j00
               - lap(L.nm1j00)
               + (-3.0 / (c * dt * dz) + 2.0 * 1j * djn / (c * dt) + 2.0 / (c * c * dt * dt)
                  + 1j * 2.0 * k0 / (c * dt)) * L.nm1j00)
        acoeff = 3.0 / (c * dt * dz) + 2.0 / (c * c * dt * dt) - 1j * 2.0 * (k0 + djn) / (c * dt)
    rhs_f = np.fft.fft2(rhs)
    dkx = 2.0 * math.pi / (geom.hi[0] - geom.lo[0])
    dky = 2.0 * math.pi / (geom.hi[1] - geom.lo[1])
    ii, jj = np.arange(Nx), np.arange(Ny)
    kx = np.where(ii < imid, dkx * ii, dkx * (ii - Nx))[None, :]
    ky = np.where(jj < jmid, dky * jj, dky * (jj - Ny))[:, None]
    den = kx * kx + ky * ky + acoeff
    inv = np.where(np.abs(den) > 0.0, 1.0 / np.where(den == 0, 1.0, den), 0.0)
    L.np1j00 = np.fft.ifft2(-rhs_f * inv)


def laser_gather(xp, yp, aabs, geom: 'Geometry', derivatives: bool):
    """doLaserGatherShapeN<order>, particles_utils/FieldGather.H:236-283 (value + centred
    derivatives taken on the grid, then gathered) and :299-330 (value only)"""
    G, n = geom.g, geom.order + 1
    x = (xp - geom.pos_offset(0)) * (1.0 / geom.dx)
    y = (yp - geom.pos_offset(1)) * (1.0 / geom.dy)
    sx, i0 = shape(geom.order, x)
    sy, j0 = shape(geom.order, y)
    dx_inv, dy_inv = 1.0 / geom.dx, 1.0 / geom.dy
    A = np.zeros_like(xp); ADx = np.zeros_like(xp); ADy = np.zeros_like(xp)
    for iy in range(n):
        for ix in range(n):
            i, j = i0 + ix + G, j0 + iy + G
            w = sx[ix] * sy[iy]
            A += w * aabs[j, i]
            if derivatives:
                ADx += w * 0.5 * dx_inv * (aabs[j, i + 1] - aabs[j, i - 1])
                ADy += w * 0.5 * dy_inv * (aabs[j + 1, i] - aabs[j - 1, i])
    return A, ADx, ADy


def _scatter(arr, jj, ii, vals, G):
    """arr[jj+G, ii+G] += vals, duplicates accumulated (Gpu::Atomic::Add semantics).
    np.bincount sums in a fixed order, so the oracle itself is deterministic."""
    ny_t, nx_t = arr.shape
    flat = (jj + G) * nx_t + (ii + G)
    arr += np.bincount(flat.ravel(), weights=vals.ravel(), minlength=arr.size).reshape(arr.shape)


def deposit_current(pl: Plasma, F: dict, geom: Geometry, pc: PhysConst, normalized: bool,
                    *, jx=None, jy=None, rho=None, chi=None, rhomjz=None, flip_charge=False,
                    aabs=None, jz=None):
    """::DepositCurrent, particles/deposition/PlasmaDepositCurrent.cpp:22-257 (jz not needed by
    the explicit solver).  Arguments jx.. are the destination arrays or None (the reference's -1).
    Returns the number of QSA-violating particles killed in this call (:197-204)."""
    charge = -pl.charge if flip_charge else pl.charge                       # :40
    invvol = 1.0 if normalized else 1.0 / (geom.dx * geom.dy * geom.dz)     # :71-73 (lev 0)
    x_off, y_off = geom.pos_offset(0), geom.pos_offset(1)
    dx_inv, dy_inv = 1.0 / geom.dx, 1.0 / geom.dy
    clightinv = 1.0 / pc.c
    charge_invvol = charge * invvol
    charge_mu0_mass_ratio = charge * pc.mu0 / pl.mass

    v = pl.valid
    psi_inv = 1.0 / pl.psi
    vx_c = pl.ux * psi_inv
    vy_c = pl.uy * psi_inv
    Aabssqp = 0.0
    if aabs is not None:                                                    # :182-187
        laser_norm = (charge / pc.q_e) * (pc.m_e / pl.mass) * (charge / pc.q_e) * (pc.m_e / pl.mass)
        Aabssqp = laser_gather(pl.x, pl.y, aabs, geom, False)[0] * laser_norm
    gamma_psi = 0.5 * ((1.0 + 0.5 * Aabssqp) * psi_inv * psi_inv + vx_c * vx_c * clightinv * clightinv
                       + vy_c * vy_c * clightinv * clightinv + 1.0)         # :190-195
    bad = v & ((gamma_psi < 0.0) | (gamma_psi > pl.max_qsa_weighting_factor) | (psi_inv < 0.0))
    n_bad = int(bad.sum())
    if n_bad:
        pl.w[bad] = 0.0
        pl.valid[bad] = False
    sel = pl.valid.copy()
    if not sel.any():
        return n_bad
    xmid = (pl.x[sel] - x_off) * dx_inv
    ymid = (pl.y[sel] - y_off) * dy_inv
    sx, i0 = shape(geom.order, xmid)
    sy, j0 = shape(geom.order, ymid)
    q_invvol = charge_invvol * pl.w[sel]
    psi_inv, vx_c, vy_c, gamma_psi = psi_inv[sel], vx_c[sel], vy_c[sel], gamma_psi[sel]
    n, G = geom.order + 1, geom.g
    ii = np.stack([i0 + ix for iy in range(n) for ix in range(n)])
    jj = np.stack([j0 + iy for iy in range(n) for ix in range(n)])
    cd = np.stack([q_invvol * sx[ix] * sy[iy] for iy in range(n) for ix in range(n)])  # :218
    if jx is not None:
        _scatter(jx, jj, ii, cd * vx_c, G)
        _scatter(jy, jj, ii, cd * vy_c, G)
    if jz is not None:                                                      # :223 (predictor-corrector)
        _scatter(jz, jj, ii, cd * (gamma_psi - 1.0) * pc.c, G)
    if rho is not None:
        _scatter(rho, jj, ii, cd * gamma_psi, G)
    if chi is not None:
        _scatter(chi, jj, ii, cd * charge_mu0_mass_ratio * psi_inv, G)
    if rhomjz is not None:
        _scatter(rhomjz, jj, ii, cd, G)
    return n_bad


def beam_deposit(bs: dict, beam: Beam, geom: Geometry, pc: PhysConst, normalized: bool,
                 *, jxb=None, jyb=None, jzb=None):
    """DepositCurrentSlice, particles/deposition/BeamDepositCurrent.cpp:21-195 (lev 0)."""
    if bs is None or bs['x'].size == 0:
        return
    if bs.get('np', bs['x'].size) != bs['x'].size:      # getNumParticles excludes slipped (:100)
        n = bs['np']
        bs = {k: (v[:n] if isinstance(v, np.ndarray) else v) for k, v in bs.items()}
    invvol = 1.0 if normalized else 1.0 / (geom.dx * geom.dy * geom.dz)     # :72-82
    x_off, y_off = geom.pos_offset(0), geom.pos_offset(1)
    clightsq = 1.0 / (pc.c * pc.c)
    ux, uy, uz = bs['ux'], bs['uy'], bs['uz']
    gaminv = 1.0 / np.sqrt(1.0 + ux * ux * clightsq + uy * uy * clightsq + uz * uz * clightsq)
    wq = beam.charge * bs['w'] * invvol
    sx, i0 = shape(geom.order, (bs['x'] - x_off) / geom.dx)
    sy, j0 = shape(geom.order, (bs['y'] - y_off) / geom.dy)
    n, G = geom.order + 1, geom.g
    ii = np.stack([i0 + ix for iy in range(n) for ix in range(n)])
    jj = np.stack([j0 + iy for iy in range(n) for ix in range(n)])
    ss = np.stack([sx[ix] * sy[iy] for iy in range(n) for ix in range(n)])
    v = bs['valid']
    ii, jj, ss = ii[:, v], jj[:, v], ss[:, v]
    if jxb is not None:
        _scatter(jxb, jj, ii, ss * (wq * ux * gaminv)[v], G)
        _scatter(jyb, jj, ii, ss * (wq * uy * gaminv)[v], G)
    if jzb is not None:
        _scatter(jzb, jj, ii, ss * (wq * uz * gaminv)[v], G)


def grid_current_deposit(jz, geom: Geometry, islice: int, peak, mean, std):
    """GridCurrent::DepositCurrentSlice, utils/GridCurrent.cpp:25-70: an analytic gaussian current
    density added to jz_beam on the valid box (cell centres plo + (i + 1/2) dx, z = plo + islice dz)"""
    G = geom.g
    z = geom.lo[2] + islice * geom.dz
    dz_ = (z - mean[2]) / std[2]
    long_f = math.exp(-0.5 * (dz_ * dz_))
    x = geom.lo[0] + (np.arange(geom.nx) + 0.5) * geom.dx
    y = geom.lo[1] + (np.arange(geom.ny) + 0.5) * geom.dy
    ddx = ((x - mean[0]) / std[0])[None, :]
    ddy = ((y - mean[1]) / std[1])[:, None]
    jz[G:-G, G:-G] += peak * np.exp(-0.5 * (ddx * ddx + ddy * ddy)) * long_f


def poisson_eigenvalues(nx, ny, dx, dy):
    """m_eigenvalue_matrix incl. DST normalisation,
    fields/fft_poisson_solver/FFTPoissonSolverDirichletFast.cpp:224-248."""
    sx = np.sin((np.arange(nx) + 1) * (math.pi / (2.0 * (nx + 1)))) ** 2
    sy = np.sin((np.arange(ny) + 1) * (math.pi / (2.0 * (ny + 1)))) ** 2
    norm_fac = 0.5 / (2 * ((nx + 1) * (ny + 1)))
    return norm_fac / (-4.0 * (sx[None, :] / (dx * dx) + sy[:, None] / (dy * dy)))


def poisson_dirichlet(rhs, eig):
    """FFTPoissonSolverDirichletFast::SolvePoissonEquation (:286-328) ==
    FFTPoissonSolverDirichletDirect (FFTW RODFT00): lhs = DST2D(DST2D(rhs) * eig).
    scipy dstn(type=1) is FFTW's unnormalised RODFT00 (y_k = 2 sum x_j sin(pi (j+1)(k+1)/(n+1)))."""
    return _dstn(_dstn(rhs, type=1) * eig, type=1)


def poisson_periodic(rhs, dx, dy):
    """FFTPoissonSolverPeriodic::SolvePoissonEquation (fields/fft_poisson_solver/FFTPoissonSolverPeriodic.cpp
    :111-149): R2C 2-D FFT, times -inv_k2, C2R, times 1/N.  inv_k2 (:69-92): half spectrum along x,
    kx = dkx i, ky = dky j (j < (Ny+1)/2) else dky (j - Ny), and ZERO wherever i == 0 or j == 0 (the whole
    kx = 0 row and ky = 0 column, not only the origin)."""
    ny, nx = rhs.shape
    dkx, dky = 2 * math.pi / (nx * dx), 2 * math.pi / (ny * dy)
    i = np.arange(nx // 2 + 1)
    j = np.arange(ny)
    kx = dkx * i
    ky = np.where(j < (ny + 1) // 2, dky * j, dky * (j - ny))
    k2 = kx[None, :] ** 2 + ky[:, None] ** 2
    inv_k2 = np.zeros_like(k2)
    m = (i[None, :] != 0) & (j[:, None] != 0)
    inv_k2[m] = 1.0 / k2[m]
    spec = np.fft.rfft2(rhs) * (-inv_k2)
    # numpy's irfft2 carries the 1/N of the reference's final copy (:136-148)
    return np.fft.irfft2(spec, s=(ny, nx))


def enforce_periodic(arrs, G, do_sum):
    """Fields::EnforcePeriodic (fields/Fields.cpp:1117-1145) for the single slice box: AMReX SumBoundary
    (do_sum: the guard cells are ADDED to their periodic images in the valid box and keep their own
    values) or FillBoundary (the guard cells are overwritten with their periodic images)."""
    for a in arrs:
        ny, nx = a.shape[0] - 2 * G, a.shape[1] - 2 * G
        v = a[G:-G, G:-G]
        if do_sum:
            src = a.copy()
            v[:, :G] += src[G:-G, nx + G:]            # (i + nx, j)
            v[:, nx - G:] += src[G:-G, :G]            # (i - nx, j)
            v[:G, :] += src[ny + G:, G:-G]            # (i, j + ny)
            v[ny - G:, :] += src[:G, G:-G]            # (i, j - ny)
            v[:G, :G] += src[ny + G:, nx + G:]        # corners
            v[:G, nx - G:] += src[ny + G:, :G]
            v[ny - G:, :G] += src[:G, nx + G:]
            v[ny - G:, nx - G:] += src[:G, :G]
        else:
            a[...] = np.pad(v, G, mode='wrap')


def _ddx(a, dx, G):
    """derivative<x> on the valid box, fields/Fields.cpp:223-235: (f[i+1]-f[i-1]) * 0.5/dx"""
    return (a[G:-G, G + 1:a.shape[1] - G + 1] - a[G:-G, G - 1:a.shape[1] - G - 1]) * (0.5 / dx)


def _ddy(a, dy, G):
    return (a[G + 1:a.shape[0] - G + 1, G:-G] - a[G - 1:a.shape[0] - G - 1, G:-G]) * (0.5 / dy)


N_MULTIPOLE = 18      # fields/OpenBoundary.H: 1 + 2 * 18 coefficients


def open_boundary_rhs(rhs, geom: Geometry, monopole: bool):
    """Fields::SetBoundaryCondition for boundary.field = Open (fields/Fields.cpp:685-738) followed by
    SetDirichletBoundaries (:628-673), offset = factor = 1 (FFTPoissonSolverDirichletFast.H:58-59).

    The free-space potential of the sources inside 95 % of the largest centred circle,
    phi(r) = dx dy / (4 pi) sum_s s ln |r - r_s|^2, is expanded about the origin: with z = x + i y
    (both scaled by 3 / box diagonal) and the moments M_k = sum_s s z_s^k,
        phi = dx dy / (4 pi) ( M_0 ln |z|^2 - 2 sum_{k=1..18} Re(M_k / z^k) / k ),
    which is what the 37 real coefficients and the polynomial table of fields/OpenBoundary.H:39-156
    spell out term by term.  Its value one cell outside every edge cell is the non-zero Dirichlet
    value, folded into the right-hand side as -value / dx^2 (Van Loan).  Ez and Bz have no monopole."""
    ny, nx = rhs.shape
    dx, dy = geom.dx, geom.dy
    off_x = 0.5 * (geom.lo[0] + geom.hi[0] - dx * (nx - 1))
    off_y = 0.5 * (geom.lo[1] + geom.hi[1] - dy * (ny - 1))          # GetPosOffset of the valid box
    scale = 3.0 / math.sqrt((geom.hi[0] - geom.lo[0]) ** 2 + (geom.hi[1] - geom.lo[1]) ** 2)
    radius = min(abs(geom.lo[0]), abs(geom.hi[0]), abs(geom.lo[1]), abs(geom.hi[1]))
    assert radius > 0.0, 'x = 0, y = 0 must be inside the box'
    cutoff_sq = (0.95 * radius * scale) ** 2
    x = ((np.arange(nx) * dx + off_x) * scale)[None, :]
    y = ((np.arange(ny) * dy + off_y) * scale)[:, None]
    inside = ~(x * x + y * y > cutoff_sq)
    zs = (x + 1j * y)[inside]
    sv = rhs[inside]
    M = [np.sum(sv * zs ** k) for k in range(N_MULTIPOLE + 1)]
    if not monopole:
        M[0] = 0.0

    def value(xd, yd):
        z = (xd + 1j * yd) * scale
        phi = np.real(M[0]) * np.log(np.abs(z) ** 2)
        for k in range(1, N_MULTIPOLE + 1):
            phi = phi - 2.0 * np.real(M[k] / z ** k) / k
        return dx * dy / (4.0 * math.pi) * phi

    out = rhs.copy()
    xi = np.arange(nx) * dx + off_x
    yj = np.arange(ny) * dy + off_y
    out[0, :] += -value(xi, (0 - 1) * dy + off_y) / (dy * dy)             # j_lo edge
    out[ny - 1, :] += -value(xi, (ny - 1 + 1) * dy + off_y) / (dy * dy)   # j_hi edge
    out[:, 0] += -value((0 - 1) * dx + off_x, yj) / (dx * dx)             # i_lo edge
    out[:, nx - 1] += -value((nx - 1 + 1) * dx + off_x, yj) / (dx * dx)   # i_hi edge
    return out


def solve_poisson_psi_ez_bz(F, geom: Geometry, pc: PhysConst, eig, open_bc=False, field_periodic=False,
                            poisson_periodic_solver=False):
    """Fields::SolvePoissonPsiExmByEypBxEzBz, fields/Fields.cpp:840-957 (lev 0; boundary.field Dirichlet,
    Open or Periodic; fields.poisson_solver FFTDirichlet* or FFTPeriodic)."""
    dx, dy, G = geom.dx, geom.dy, geom.g
    T = lambda n: F[('This', n)]
    bc = (lambda r, mono: open_boundary_rhs(r, geom, mono)) if open_bc else (lambda r, mono: r)
    if field_periodic:                                                           # :859-861
        enforce_periodic([T('jx'), T('jy'), T('rhomjz')], G, True)
    solve = (lambda r, e: poisson_periodic(r, dx, dy)) if poisson_periodic_solver else poisson_dirichlet
    T('Psi')[G:-G, G:-G] = solve(bc((-1.0 / pc.ep0) * T('rhomjz')[G:-G, G:-G], True), eig)
    f = 1.0 / (pc.ep0 * pc.c)
    T('Ez')[G:-G, G:-G] = solve(bc(f * _ddx(T('jx'), dx, G) + f * _ddy(T('jy'), dy, G), False), eig)
    T('Bz')[G:-G, G:-G] = solve(bc(pc.mu0 * _ddy(T('jx'), dy, G)
                                               + (-pc.mu0) * _ddx(T('jy'), dx, G), False), eig)
    if field_periodic:                                                           # :920-922
        enforce_periodic([T('Psi'), T('Ez'), T('Bz')], G, False)
    # ExmBy / EypBx on the box grown by g-1, i.e. everything but the outermost ring (:931-956)
    psi = T('Psi')
    ny_t, nx_t = psi.shape
    s = (slice(1, ny_t - 1), slice(1, nx_t - 1))
    T('ExmBy')[s] = -(psi[1:ny_t - 1, 2:nx_t] - psi[1:ny_t - 1, 0:nx_t - 2]) * (0.5 / dx)
    T('EypBx')[s] = -(psi[2:ny_t, 1:nx_t - 1] - psi[0:ny_t - 2, 1:nx_t - 1]) * (0.5 / dy)


def init_sxsy_with_beam(F, geom: Geometry, pc: PhysConst):
    """Hipace::InitializeSxSyWithBeam, Hipace.cpp:744-790 (valid box)."""
    dx, dy, dz, G = geom.dx, geom.dy, geom.dz, geom.g
    jzb = F[('This', 'jz_beam')]
    ny_t, nx_t = jzb.shape
    v = (slice(G, -G), slice(G, -G))
    dx_jzb = (jzb[G:-G, G + 1:nx_t - G + 1] - jzb[G:-G, G - 1:nx_t - G - 1]) / (2.0 * dx)
    dy_jzb = (jzb[G + 1:ny_t - G + 1, G:-G] - jzb[G - 1:ny_t - G - 1, G:-G]) / (2.0 * dy)
    dz_jxb = (F[('Previous', 'jx_beam')][v] - F[('Next', 'jx_beam')][v]) / (2.0 * dz)
    dz_jyb = (F[('Previous', 'jy_beam')][v] - F[('Next', 'jy_beam')][v]) / (2.0 * dz)
    F[('This', 'Sy')][v] = pc.mu0 * (-dy_jzb + dz_jyb)
    F[('This', 'Sx')][v] = -pc.mu0 * (-dx_jzb + dz_jxb)


def explicit_deposition(pl: Plasma, F, geom: Geometry, pc: PhysConst, normalized: bool, aabs=None):
    """::ExplicitDeposition, particles/deposition/ExplicitDeposition.cpp:20-263
    (any depos_order 0..3 / derivative_type 0..2; laser terms :167-175, :211-226, :234, :250)."""
    G, nst = geom.g, geom.order + geom.dtype + 1
    sel = pl.valid
    if not sel.any():
        return
    invvol = 1.0 if normalized else 1.0 / (geom.dx * geom.dy * geom.dz)
    x_off, y_off = geom.pos_offset(0), geom.pos_offset(1)
    dx_inv, dy_inv = 1.0 / geom.dx, 1.0 / geom.dy
    clight_inv = 1.0 / pc.c
    a_clight = pc.c
    charge_invvol_mu0 = pl.charge * invvol * pc.mu0
    q_mass_ratio = pl.charge / pl.mass

    psi_inv = 1.0 / pl.psi[sel]
    vx = pl.ux[sel] * psi_inv * clight_inv
    vy = pl.uy[sel] * psi_inv * clight_inv
    cdm = charge_invvol_mu0 * pl.w[sel]
    xmid = (pl.x[sel] - x_off) * dx_inv
    ymid = (pl.y[sel] - y_off) * dy_inv
    Aabssqp = 0.0
    laser_fac = (pc.m_e / pc.q_e) * (pc.m_e / pc.q_e)                         # :55
    if aabs is not None:
        Aabssqp = laser_gather(pl.x[sel], pl.y[sel], aabs, geom, False)[0] \
            * (laser_fac * q_mass_ratio * q_mass_ratio)
    gamma_psi = 0.5 * ((1.0 + 0.5 * Aabssqp) * psi_inv * psi_inv + vx * vx + vy * vy + 1.0)   # :177-182
    sx, dsx, i0 = dshape(geom.dtype, geom.order, xmid)
    sy, dsy, j0 = dshape(geom.dtype, geom.order, ymid)
    Bz, Ez = F[('This', 'Bz')], F[('This', 'Ez')]
    ExmBy, EypBx = F[('This', 'ExmBy')], F[('This', 'EypBx')]
    Sy, Sx = F[('This', 'Sy')], F[('This', 'Sx')]
    ii_l, jj_l, sy_l, sx_l = [], [], [], []
    for iy in range(nst):
        for ix in range(nst):
            if geom.dtype == 2 and ix in (0, nst - 1) and iy in (0, nst - 1):
                continue                                                      # :193-198
            i, j = i0 + ix, j0 + iy
            shx, shdx, shy, shdy = sx[ix], dsx[ix], sy[iy], dsy[iy]
            Bz_v, Ez_v = Bz[j + G, i + G], Ez[j + G, i + G]
            ExmBy_v, EypBx_v = ExmBy[j + G, i + G], EypBx[j + G, i + G]
            ADx = ADy = 0.0
            if aabs is not None:                                               # :215-226
                nz_w = (shx * shy) != 0.0
                ic, jc = np.where(nz_w, i, 0) + G, np.where(nz_w, j, 0) + G    # "avoid going outside"
                ADx = np.where(nz_w, (aabs[jc, ic + 1] - aabs[jc, ic - 1]) * 0.5 * dx_inv * laser_fac * a_clight, 0.0)
                ADy = np.where(nz_w, (aabs[jc + 1, ic] - aabs[jc - 1, ic]) * 0.5 * dy_inv * laser_fac * a_clight, 0.0)
            val_sy = cdm * (                                                   # :228-242
                - shx * shy * (
                    - Bz_v * vx
                    + (Ez_v * vy
                       + ExmBy_v * (- vx * vy)
                       + EypBx_v * (gamma_psi - vy * vy)) * clight_inv
                    - 0.25 * ADy * q_mass_ratio * psi_inv
                ) * q_mass_ratio * psi_inv
                + (- shdx * shy * dx_inv * (- vx * vy)
                   - shx * shdy * dy_inv * (gamma_psi - vy * vy - 1.0)) * a_clight)
            val_sx = cdm * (                                                   # :244-258
                + shx * shy * (
                    + Bz_v * vy
                    + (Ez_v * vx
                       + ExmBy_v * (gamma_psi - vx * vx)
                       + EypBx_v * (- vx * vy)) * clight_inv
                    - 0.25 * ADx * q_mass_ratio * psi_inv
                ) * q_mass_ratio * psi_inv
                + (+ shdx * shy * dx_inv * (gamma_psi - vx * vx - 1.0)
                   + shx * shdy * dy_inv * (- vx * vy)) * a_clight)
            ii_l.append(i); jj_l.append(j); sy_l.append(val_sy); sx_l.append(val_sx)
    ii, jj = np.stack(ii_l), np.stack(jj_l)
    _scatter(Sy, jj, ii, np.stack(sy_l), G)
    _scatter(Sx, jj, ii, np.stack(sx_l), G)


# ---- hpmg (mg_solver/HpMultiGrid.cpp), system type 1 ---------------------------------------

class MultiGrid1:
    """hpmg::MultiGrid with system_type 1: solves  lap(phi) - acf*phi = rhs  for two components
    sharing acf, Dirichlet 0 (cell-centred: at the cell face, :163-182; nodal: at nodes 0, n+1).

    Level arrays are indexed [j, i] over m_domain[ilev] (:1054-1072): cell-centred for even n
    (0..n-1), nodal for odd n (nodes 0..n+1 where 0 and n+1 are boundary nodes held at 0).
    The CPU (gsrb_cached :594-740) and CUDA (gsrb_shared :413-590, bottomsolve_gpu :854-1033)
    paths compute the same values as whole-array red-black sweeps, which is what we do here.
    """

    def __init__(self, dx, dy, nx, ny):
        assert nx % 2 == ny % 2, 'HpMultiGrid.cpp:1051-1052'
        self.dx, self.dy = dx, dy
        self.cc = (nx % 2 == 0)
        if self.cc:
            sizes = [(ny, nx)]
            minw = 2
            while True:                                 # coarsenable(2, min_width) :1065-1072
                h, w = sizes[-1]
                if h % 2 == 0 and w % 2 == 0 and h >= 2 * minw and w >= 2 * minw:
                    sizes.append((h // 2, w // 2))
                else:
                    break
        else:
            # nodal box 0..n+1 (n+2 points); coarsenable if hi even and npts >= 2*4
            sizes = [(ny + 2, nx + 2)]
            while True:
                h, w = sizes[-1]
                if (h - 1) % 2 == 0 and (w - 1) % 2 == 0 and h >= 8 and w >= 8:
                    sizes.append(((h - 1) // 2 + 1, (w - 1) // 2 + 1))
                else:
                    break
        self.sizes = sizes
        self.nlev = len(sizes)
        self.acf = [np.zeros(s) for s in sizes]
        self.res = [None] + [np.zeros((2,) + s) for s in sizes[1:]]
        self.cor = [np.zeros((2,) + s) for s in sizes]
        self.rescor = [np.zeros((2,) + s) for s in sizes]
        self.n_vcycles_last = 0

    # -- helpers -------------------------------------------------------------------------
    def _valid(self, a):
        return a if self.cc else a[..., 1:-1, 1:-1]

    def _masks(self, shape):
        ny, nx = shape
        j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing='ij')
        return (i + j)

    def _lap_offdiag(self, phi, facx, facy):
        """sum of neighbour terms as in gs1 (:265-292) and the matching diagonal c0 modifier."""
        ny, nx = phi.shape[-2:]
        p = np.zeros(phi.shape[:-2] + (ny + 2, nx + 2))
        p[..., 1:-1, 1:-1] = phi
        lapx = facx * (p[..., 1:-1, :-2] + p[..., 1:-1, 2:])
        lapy = facy * (p[..., :-2, 1:-1] + p[..., 2:, 1:-1])
        if self.cc:
            lapx[..., :, 0] = facx * (4. / 3.) * phi[..., :, 1]
            lapx[..., :, -1] = facx * (4. / 3.) * phi[..., :, -2]
            lapy[..., 0, :] = facy * (4. / 3.) * phi[..., 1, :]
            lapy[..., -1, :] = facy * (4. / 3.) * phi[..., -2, :]
        return lapx + lapy

    def _c0(self, acf, facx, facy):
        c0 = -(acf + 2.0 * (facx + facy))
        if self.cc:
            c0 = c0.copy()
            c0[:, 0] -= 2.0 * facx
            c0[:, -1] -= 2.0 * facx
            c0[0, :] -= 2.0 * facy
            c0[-1, :] -= 2.0 * facy
        return c0

    def gsrb(self, phi, rhs, acf, facx, facy, nsweeps, first_color=0):
        """nsweeps red-black half-sweeps, colour (i+j+icolor)%2==0 (:367-404, :521-548).
        phi[2, ny, nx] over the level box; for nodal only the interior is updated."""
        if self.cc:
            par = self._masks(phi.shape[-2:])
            c0_inv = 1.0 / self._c0(acf, facx, facy)
            for ic in range(first_color, first_color + nsweeps):
                m = ((par + ic) % 2 == 0)
                lap = self._lap_offdiag(phi, facx, facy)
                new = (rhs - lap) * c0_inv
                phi[:, m] = new[:, m]
        else:
            ny, nx = phi.shape[-2:]
            j, i = np.meshgrid(np.arange(1, ny - 1), np.arange(1, nx - 1), indexing='ij')
            par = i + j
            c0_inv = 1.0 / self._c0(acf[1:-1, 1:-1], facx, facy)
            for ic in range(first_color, first_color + nsweeps):
                m = ((par + ic) % 2 == 0)
                lap = (facx * (phi[:, 1:-1, :-2] + phi[:, 1:-1, 2:])
                       + facy * (phi[:, :-2, 1:-1] + phi[:, 2:, 1:-1]))
                new = (rhs[:, 1:-1, 1:-1] - lap) * c0_inv
                inner = phi[:, 1:-1, 1:-1]
                inner[:, m] = new[:, m]
        return phi

    def residual(self, phi, rhs, acf, facx, facy):
        """residual1 (:184-190): rhs + acf*phi - laplacian(phi), laplacian as :163-182."""
        out = np.zeros_like(phi)
        if self.cc:
            lap = -2.0 * (facx + facy) * phi
            lx = self._lap_offdiag_x(phi, facx)
            ly = self._lap_offdiag_y(phi, facy)
            lap = lap + lx
            lap = lap + ly
            out[:] = rhs + acf * phi - lap
        else:
            inner = phi[:, 1:-1, 1:-1]
            lap = -2.0 * (facx + facy) * inner
            lap = lap + facx * (phi[:, 1:-1, :-2] + phi[:, 1:-1, 2:])
            lap = lap + facy * (phi[:, :-2, 1:-1] + phi[:, 2:, 1:-1])
            out[:, 1:-1, 1:-1] = rhs[:, 1:-1, 1:-1] + acf[1:-1, 1:-1] * inner - lap
        return out

    def _lap_offdiag_x(self, phi, facx):
        lx = np.empty_like(phi)
        lx[..., :, 1:-1] = facx * (phi[..., :, :-2] + phi[..., :, 2:])
        lx[..., :, 0] = facx * ((4. / 3.) * phi[..., :, 1] - 2.0 * phi[..., :, 0])
        lx[..., :, -1] = facx * ((4. / 3.) * phi[..., :, -2] - 2.0 * phi[..., :, -1])
        return lx

    def _lap_offdiag_y(self, phi, facy):
        ly = np.empty_like(phi)
        ly[..., 1:-1, :] = facy * (phi[..., :-2, :] + phi[..., 2:, :])
        ly[..., 0, :] = facy * ((4. / 3.) * phi[..., 1, :] - 2.0 * phi[..., 0, :])
        ly[..., -1, :] = facy * ((4. / 3.) * phi[..., -2, :] - 2.0 * phi[..., -1, :])
        return ly

    def restrict(self, fine):
        """restrict_cc (:29-37) / restrict_nd (:39-52)"""
        if self.cc:
            return 0.25 * (fine[..., 0::2, 0::2] + fine[..., 0::2, 1::2]
                           + fine[..., 1::2, 0::2] + fine[..., 1::2, 1::2])
        nyc = (fine.shape[-2] - 1) // 2 + 1
        nxc = (fine.shape[-1] - 1) // 2 + 1
        crse = np.zeros(fine.shape[:-2] + (nyc, nxc))
        c = lambda dj, di: fine[..., 2 + dj:fine.shape[-2] - 2 + dj + 1:2,
                                2 + di:fine.shape[-1] - 2 + di + 1:2]
        crse[..., 1:-1, 1:-1] = (1. / 16.) * (
            c(-1, -1) + 2. * c(-1, 0) + c(-1, 1)
            + 2. * c(0, -1) + 4. * c(0, 0) + 2. * c(0, 1)
            + c(1, -1) + 2. * c(1, 0) + c(1, 1))
        return crse

    def interp_add(self, fine, crse):
        """interpcpy_cc / interpcpy_nd (:88-121): returns fine + I(crse)"""
        if self.cc:
            return fine + np.repeat(np.repeat(crse, 2, axis=-2), 2, axis=-1)
        out = fine.copy()
        ny, nx = fine.shape[-2:]
        I = np.zeros_like(fine)
        I[..., 0::2, 0::2] = crse
        I[..., 1::2, 0::2] = (crse[..., :-1, :] + crse[..., 1:, :]) * 0.5
        I[..., 0::2, 1::2] = (crse[..., :, :-1] + crse[..., :, 1:]) * 0.5
        I[..., 1::2, 1::2] = (crse[..., :-1, :-1] + crse[..., 1:, :-1]
                              + crse[..., :-1, 1:] + crse[..., 1:, 1:]) * 0.25
        out[..., 1:-1, 1:-1] += I[..., 1:-1, 1:-1]
        return out

    # -- solve1 ----------------------------------------------------------------------------
    def solve1(self, sol, rhs, acf, tol_rel=1e-4, tol_abs=np.finfo(float).tiny, maxiter=200):
        """solve1 (:1169-1190) + solve_doit (:1307-1427) + vcycle (:1429-1512) +
        bottomsolve (:1514-1594).  sol[2,ny,nx], rhs[2,ny,nx], acf[ny,nx] are *valid-box*
        arrays (the caller strips the guard cells = center_box, HpMultiGrid.H:168-175).
        sol is the initial guess and is overwritten."""
        dx, dy = self.dx, self.dy

        def emb(a):      # valid-box array -> level-0 box
            if self.cc:
                return a
            out = np.zeros(a.shape[:-2] + self.sizes[0])
            out[..., 1:-1, 1:-1] = a
            return out
        self.acf[0] = emb(np.array(acf, dtype=float))
        for l in range(1, self.nlev):                        # average_down_acoef :1640-1700
            self.acf[l] = self.restrict(self.acf[l - 1])
        rhs0 = emb(rhs)
        fac = lambda l: (1.0 / ((dx * (1 << l)) ** 2), 1.0 / ((dy * (1 << l)) ** 2))

        def gsrb4_res(phi_in, r, l, zero_init, do_res):
            fx, fy = fac(l)
            phi = np.zeros_like(r) if zero_init else phi_in.copy()
            self.gsrb(phi, r, self.acf[l], fx, fy, 4)
            rr = self.residual(phi, r, self.acf[l], fx, fy) if do_res else None
            return phi, rr

        sol0 = emb(sol)
        self.cor[0], self.rescor[0] = gsrb4_res(sol0, rhs0, 0, False, True)     # :1326-1327
        resnorm0 = np.abs(self._valid(self.rescor[0])).max()
        rhsnorm0 = np.abs(self._valid(rhs0)).max()
        max_norm = max(rhsnorm0, resnorm0)
        res_target = max(tol_abs, max(tol_rel, 1e-16) * max_norm)               # :1361
        self.n_vcycles_last = 0
        if resnorm0 > res_target:
            converged = False
            for it in range(maxiter):
                self._vcycle(sol0, rhs0, gsrb4_res, fac)
                self.n_vcycles_last = it + 1
                norminf = np.abs(self.rescor[0]).max()
                if norminf <= res_target:
                    converged = True
                    break
                if norminf > 1e20 * max_norm:
                    raise RuntimeError('hpmg failing so lets stop here')
            if not converged:
                raise RuntimeError('hpmg failed')
        sol[...] = self._valid(self.cor[0])                                    # :1419-1426
        return sol

    def _vcycle(self, sol0, rhs0, gsrb4_res, fac):
        nl = self.nlev
        for l in range(0, nl - 1):
            if l > 0:
                self.cor[l], self.rescor[l] = gsrb4_res(None, self.res[l], l, True, True)
            self.res[l + 1] = self.restrict(self.rescor[l])
        # bottom: 16 sweeps (numsweeps = max(16, (len.max()+1)/2*2), :1587; coarsest len <= 5)
        l = nl - 1
        fx, fy = fac(l)
        nsw = max(16, (max(self.sizes[l]) + 1) // 2 * 2)
        self.cor[l] = np.zeros_like(self.res[l])
        self.gsrb(self.cor[l], self.res[l], self.acf[l], fx, fy, nsw)
        for l in range(nl - 2, -1, -1):
            self.rescor[l] = self.interp_add(self.cor[l], self.cor[l + 1])
            if l == 0:
                new, _ = gsrb4_res(self.rescor[0], rhs0, 0, False, False)
                sol0[...] = new
            else:
                self.cor[l], _ = gsrb4_res(self.rescor[l], self.res[l], l, False, False)
        self.cor[0], self.rescor[0] = gsrb4_res(sol0, rhs0, 0, False, True)     # :1501-1503


class MultiGrid2(MultiGrid1):
    """hpmg::MultiGrid with system_type 2 (the laser envelope solver, laser/MultiLaser.cpp:598-606):
    the complex system  lap(A) - (a_r + i a_i) A = rhs  as two coupled real components; gs2
    (mg_solver/HpMultiGrid.cpp:296-334), residual2r / residual2i (:192-208); everything else (V-cycle,
    transfer operators, stopping rule) is shared with type 1.  acf arrays are [2, ny, nx] = (a_r, a_i).
    Not pinned by a golden of its own (the reference checksums only its fft-solver run): validated
    against the pinned fft restatement to the multigrid tolerance (tests/test_oracle_golden.py)."""

    def gsrb(self, phi, rhs, acf, facx, facy, nsweeps, first_color=0):
        ar, ai = acf[0], acf[1]
        if self.cc:
            par = self._masks(phi.shape[-2:])
            c0 = self._c0(np.zeros_like(ar), facx, facy)          # -2 (facx + facy) + boundary terms
            sl = (slice(None), slice(None))
        else:
            ny, nx = phi.shape[-2:]
            j, i = np.meshgrid(np.arange(1, ny - 1), np.arange(1, nx - 1), indexing='ij')
            par = i + j
            ar, ai = ar[1:-1, 1:-1], ai[1:-1, 1:-1]
            c0 = -2.0 * (facx + facy) + np.zeros_like(ar)
            sl = (slice(1, -1), slice(1, -1))
        c_r, c_i = c0 - ar, -ai
        cmag = 1.0 / (c_r * c_r + c_i * c_i)
        c_r, c_i = c_r * cmag, c_i * cmag
        for ic in range(first_color, first_color + nsweeps):
            m = ((par + ic) % 2 == 0)
            if self.cc:
                lap = self._lap_offdiag(phi, facx, facy)
            else:
                lap = (facx * (phi[:, 1:-1, :-2] + phi[:, 1:-1, 2:])
                       + facy * (phi[:, :-2, 1:-1] + phi[:, 2:, 1:-1]))
            dr, di = rhs[0][sl] - lap[0], rhs[1][sl] - lap[1]
            new_r = dr * c_r + di * c_i
            new_i = di * c_r - dr * c_i
            v = phi[(slice(None),) + sl]
            v[0][m] = new_r[m]
            v[1][m] = new_i[m]
        return phi

    def residual(self, phi, rhs, acf, facx, facy):
        out = np.zeros_like(phi)
        if self.cc:
            lap = -2.0 * (facx + facy) * phi
            lap = lap + self._lap_offdiag_x(phi, facx)
            lap = lap + self._lap_offdiag_y(phi, facy)
            out[0] = rhs[0] + acf[0] * phi[0] - acf[1] * phi[1] - lap[0]
            out[1] = rhs[1] + acf[1] * phi[0] + acf[0] * phi[1] - lap[1]
        else:
            inner = phi[:, 1:-1, 1:-1]
            lap = -2.0 * (facx + facy) * inner
            lap = lap + facx * (phi[:, 1:-1, :-2] + phi[:, 1:-1, 2:])
            lap = lap + facy * (phi[:, :-2, 1:-1] + phi[:, 2:, 1:-1])
            ar, ai = acf[0][1:-1, 1:-1], acf[1][1:-1, 1:-1]
            out[0, 1:-1, 1:-1] = rhs[0, 1:-1, 1:-1] + ar * inner[0] - ai * inner[1] - lap[0]
            out[1, 1:-1, 1:-1] = rhs[1, 1:-1, 1:-1] + ai * inner[0] + ar * inner[1] - lap[1]
        return out

    def solve2(self, sol, rhs, acf_r, acf_i, tol_rel=1e-4, tol_abs=0.0, maxiter=200):
        """solve2 (:1192-1296): acf_r / acf_i may be scalars or valid-box arrays"""
        shape = sol.shape[-2:]
        acf = np.stack([np.broadcast_to(np.asarray(acf_r, dtype=float), shape),
                        np.broadcast_to(np.asarray(acf_i, dtype=float), shape)])
        return self.solve1(sol, rhs, acf, tol_rel, tol_abs, maxiter)


def laser_advance_mg(L: 'LaserSlices', chi, geom: 'Geometry', pc: 'PhysConst', lambda0, dt, step, mg,
                     use_phase=True, do_avg_rhs=True, tol_rel=1e-4, tol_abs=0.0):
    """MultiLaser::AdvanceSliceMG, laser/MultiLaser.cpp:429-607 (the default laser solver): the same
    discretisation as the fft variant with chi treated per cell in the operator (do_avg_rhs) and
    the Helmholtz system solved by hpmg type 2 from the previous slice's A^{n+1} as initial guess"""
    dx, dy, dz, c = geom.dx, geom.dy, geom.dz, pc.c
    k0 = 2.0 * math.pi / lambda0
    Ny, Nx = L.n00j00.shape
    imid, jmid = (Nx + 1) // 2, (Ny + 1) // 2
    tj00 = tjp1 = tjp2 = 0.0
    if use_phase:
        keep_x = [imid - 1, imid] if Nx % 2 == 0 else [imid]
        keep_y = [jmid - 1, jmid] if Ny % 2 == 0 else [jmid]
        ax = lambda a: a[np.ix_(keep_y, keep_x)].sum()
        h0, h1, h2 = ax(L.n00j00), ax(L.n00jp1), ax(L.n00jp2)
        tj00, tjp1, tjp2 = (math.atan2(h.imag, h.real) for h in (h0, h1, h2))
    dt1, dt2 = tj00 - tjp1, tjp1 - tjp2
    if dt1 < -1.5 * math.pi: dt1 += 2.0 * math.pi
    if dt1 > 1.5 * math.pi: dt1 -= 2.0 * math.pi
    if dt2 < -1.5 * math.pi: dt2 += 2.0 * math.pi
    if dt2 > 1.5 * math.pi: dt2 -= 2.0 * math.pi
    exp1 = np.exp(1j * (tj00 - tjp1))
    exp2 = np.exp(1j * (tj00 - tjp2))
    djn = (-3.0 * dt1 + dt2) / (2.0 * dz)
    if step == 0:
        acoeff_real = 6.0 / (c * dt * dz)
        acoeff_imag = -4.0 * (k0 + djn) / (c * dt)
    else:
        acoeff_real = 3.0 / (c * dt * dz) + 2.0 / (c * c * dt * dt)
        acoeff_imag = -2.0 * (k0 + djn) / (c * dt)

    def lap(a):
        out = np.zeros_like(a)
        out[1:-1, 1:-1] = (a[1:-1, 2:] + a[1:-1, :-2] - 2.0 * a[1:-1, 1:-1]) / (dx * dx) \
            + (a[2:, 1:-1] + a[:-2, 1:-1] - 2.0 * a[1:-1, 1:-1]) / (dy * dy)
        return out
    an00j00 = L.n00j00
    if step == 0:
        rhs = (8.0 / (c * dt * dz) * (-L.np1jp1 + L.n00jp1) * exp1
               + 2.0 / (c * dt * dz) * (L.np1jp2 - L.n00jp2) * exp2
               - lap(L.n00j00)
               + (-6.0 / (c * dt * dz) + 4.0 * 1j * djn / (c * dt) + 1j * 4.0 * k0 / (c * dt)) * an00j00)
        rhs = rhs + (chi * an00j00 if do_avg_rhs else chi * an00j00 * 2.0)
    else:
        rhs = (4.0 / (c * dt * dz) * (-L.np1jp1 + L.nm1jp1) * exp1
               + 1.0 / (c * dt * dz) * (L.np1jp2 - L.nm1jp2) * exp2
               - 4.0 / (c * c * dt * dt) * an00j00
               - lap(L.nm1j00)
               + (-3.0 / (c * dt * dz) + 2.0 * 1j * djn / (c * dt) + 2.0 / (c * c * dt * dt)
                  + 1j * 2.0 * k0 / (c * dt)) * L.nm1j00)
        rhs = rhs + (chi * L.nm1j00 if do_avg_rhs else chi * an00j00 * 2.0)
    acf_r = acoeff_real + chi if do_avg_rhs else acoeff_real
    sol = np.stack([L.np1j00.real, L.np1j00.imag]).copy()          # initial guess: what np1j00 holds
    mg.solve2(sol, np.stack([rhs.real, rhs.imag]), acf_r, acoeff_imag, tol_rel, tol_abs, 200)
    L.np1j00 = sol[0] + 1j * sol[1]


# ---- gather + push ------------------------------------------------------------------------

def _momentum_push(ux, uy, psi_inv, ExmBy, EypBx, Ez, Bx_c, By_c, Bz, clight_inv, qmc,
                   A=0.0, ADx=0.0, ADy=0.0):
    """PlasmaMomentumPush<Real>, particles/pusher/PushPlasmaParticles.H:39-75
    (A, ADx, ADy = Aabssq_norm, AabssqDx_norm, AabssqDy_norm; 0 without a laser)."""
    gamma_psi = 0.5 * psi_inv * psi_inv * (1.0 + A + ux * ux * (clight_inv * clight_inv)
                                           + uy * uy * (clight_inv * clight_inv)) + 0.5
    dz_ux = qmc * (gamma_psi * ExmBy + By_c + (uy * Bz) * psi_inv) - ADx * psi_inv
    dz_uy = qmc * (gamma_psi * EypBx - Bx_c - (ux * Bz) * psi_inv) - ADy * psi_inv
    dz_psi = qmc * clight_inv * ((ux * ExmBy + uy * EypBx) * clight_inv * psi_inv - Ez)
    return dz_ux, dz_uy, dz_psi


def _momentum_push_dual(ux, uxe, uy, uye, pi, pie, ExmBy, EypBx, Ez, Bx_c, By_c, Bz,
                        clight_inv, qmc, A=0.0, ADx=0.0, ADy=0.0):
    """PlasmaMomentumPush<DualNumber> epsilon parts; arithmetic follows utils/DualNumbers.H:13-43
    operator by operator (Real*Dual promotes the Real to a Dual with epsilon 0)."""
    c2 = clight_inv * clight_inv

    def mul(a, ae, b, be):
        return a * b, ae * b + a * be

    def add(a, ae, b, be):
        return a + b, ae + be

    def sub(a, ae, b, be):
        return a - b, ae - be
    z = 0.0
    # gamma_psi = 0.5*psi_inv*psi_inv*(1 + Aabssq + ux*ux*c2 + uy*uy*c2) + 0.5
    t, te = mul(0.5, z, pi, pie)
    t, te = mul(t, te, pi, pie)
    uxx, uxxe = mul(ux, uxe, ux, uxe)
    uxx, uxxe = mul(uxx, uxxe, c2, z)
    uyy, uyye = mul(uy, uye, uy, uye)
    uyy, uyye = mul(uyy, uyye, c2, z)
    s, se = add(1.0 + A, z, uxx, uxxe)
    s, se = add(s, se, uyy, uyye)
    gp, gpe = mul(t, te, s, se)
    gp, gpe = add(gp, gpe, 0.5, z)
    # dz_ux = qmc*(gamma_psi*ExmBy + By_c + (uy*Bz)*psi_inv) - 0*psi_inv
    a, ae = mul(gp, gpe, ExmBy, z)
    a, ae = add(a, ae, By_c, z)
    b, be = mul(uy, uye, Bz, z)
    b, be = mul(b, be, pi, pie)
    a, ae = add(a, ae, b, be)
    dux, duxe = mul(qmc, z, a, ae)
    b, be = mul(ADx, z, pi, pie)
    dux, duxe = sub(dux, duxe, b, be)
    # dz_uy = qmc*(gamma_psi*EypBx - Bx_c - (ux*Bz)*psi_inv)
    a, ae = mul(gp, gpe, EypBx, z)
    a, ae = sub(a, ae, Bx_c, z)
    b, be = mul(ux, uxe, Bz, z)
    b, be = mul(b, be, pi, pie)
    a, ae = sub(a, ae, b, be)
    duy, duye = mul(qmc, z, a, ae)
    b, be = mul(ADy, z, pi, pie)
    duy, duye = sub(duy, duye, b, be)
    # dz_psi = qmc*clight_inv*((ux*ExmBy + uy*EypBx)*clight_inv*psi_inv - Ez)
    a, ae = mul(ux, uxe, ExmBy, z)
    b, be = mul(uy, uye, EypBx, z)
    a, ae = add(a, ae, b, be)
    a, ae = mul(a, ae, clight_inv, z)
    a, ae = mul(a, ae, pi, pie)
    a, ae = sub(a, ae, Ez, z)
    dps, dpse = mul(qmc * clight_inv, z, a, ae)
    return duxe, duye, dpse


def gather_fields(xp, yp, F, geom: Geometry):
    """doGatherShapeN<order>, particles/particles_utils/FieldGather.H:45-96 (always the nodal
    derivative shapes, whatever hipace.depos_derivative_type says)"""
    G, nst = geom.g, geom.order + 2
    x_off, y_off = geom.pos_offset(0), geom.pos_offset(1)
    dx_inv, dy_inv = 1.0 / geom.dx, 1.0 / geom.dy
    x = (xp - x_off) * dx_inv
    y = (yp - y_off) * dy_inv
    sx, dsx, i0 = dshape(1, geom.order, x)
    sy, dsy, j0 = dshape(1, geom.order, y)
    Psi, Ez, Bx, By, Bz = (F[('This', n)] for n in ('Psi', 'Ez', 'Bx', 'By', 'Bz'))
    ExmByp = np.zeros_like(xp); EypBxp = np.zeros_like(xp); Ezp = np.zeros_like(xp)
    Bxp = np.zeros_like(xp); Byp = np.zeros_like(xp); Bzp = np.zeros_like(xp)
    for iy in range(nst):
        for ix in range(nst):
            i, j = i0 + ix + G, j0 + iy + G
            psi_v = Psi[j, i]
            ExmByp += (dsx[ix] * sy[iy]) * psi_v * dx_inv
            EypBxp += (sx[ix] * dsy[iy]) * psi_v * dy_inv
            w = sx[ix] * sy[iy]
            Ezp += w * Ez[j, i]
            Bxp += w * Bx[j, i]
            Byp += w * By[j, i]
            Bzp += w * Bz[j, i]
    return ExmByp, EypBxp, Ezp, Bxp, Byp, Bzp


def enforce_bc(x, y, ux, uy, lo, hi, kind):
    """EnforceBC, particles/pusher/GetAndSetPosition.H:29-99.  Returns 'invalid' mask."""
    out = (x < lo[0]) | (y < lo[1]) | (x > hi[0]) | (y > hi[1])
    invalid = np.zeros_like(out)
    if not out.any():
        return invalid
    len_x, len_y = hi[0] - lo[0], hi[1] - lo[1]
    if kind == 'Periodic':
        xn = np.fmod(x[out] - lo[0], len_x); xn = np.where(xn < 0, xn + len_x, xn) + lo[0]
        yn = np.fmod(y[out] - lo[1], len_y); yn = np.where(yn < 0, yn + len_y, yn) + lo[1]
        x[out], y[out] = xn, yn
    elif kind == 'Reflecting':
        xn = np.fmod(x[out] - lo[0], 2 * len_x); xn = np.where(xn < 0, xn + 2 * len_x, xn) + lo[0]
        fx = xn > hi[0]
        xn = np.where(fx, 2 * hi[0] - xn, xn)
        yn = np.fmod(y[out] - lo[1], 2 * len_y); yn = np.where(yn < 0, yn + 2 * len_y, yn) + lo[1]
        fy = yn > hi[1]
        yn = np.where(fy, 2 * hi[1] - yn, yn)
        x[out], y[out] = xn, yn
        uxo, uyo = ux[out], uy[out]
        ux[out] = np.where(fx, -uxo, uxo)
        uy[out] = np.where(fy, -uyo, uyo)
    else:  # Absorbing
        invalid = out
    return invalid


def advance_plasma_particles(pl: Plasma, F, geom: Geometry, pc: PhysConst, bc_kind: str,
                             bc_lo, bc_hi, temp_slice=False, aabs=None):
    """AdvancePlasmaParticles, particles/pusher/PlasmaParticleAdvance.cpp:29-217 (leap-frog,
    lev 0, no laser, no ionization)."""
    sel = np.nonzero(pl.valid)[0]
    if sel.size == 0:
        return
    clight, clight_inv = pc.c, 1.0 / pc.c
    qmc = pl.charge / (pl.mass * pc.c)
    dz = geom.dz / pl.n_subcycles
    nsub = 4
    sdz = dz / nsub
    for _ in range(pl.n_subcycles):
        if sel.size == 0:
            break
        xp = pl.x_prev[sel].copy()
        yp = pl.y_prev[sel].copy()
        ExmByp, EypBxp, Ezp, Bxp, Byp, Bzp = gather_fields(xp, yp, F, geom)
        Bxp = Bxp * clight
        Byp = Byp * clight
        Ap = ADxp = ADyp = 0.0
        if aabs is not None:                                                     # :123-133
            laser_norm = (pl.charge / pc.q_e) * (pc.m_e / pl.mass) * (pl.charge / pc.q_e) * (pc.m_e / pl.mass)
            Ap, ADxp, ADyp = laser_gather(xp, yp, aabs, geom, True)
            Ap = Ap * (0.5 * laser_norm)
            ADxp = ADxp * (0.25 * clight * laser_norm)
            ADyp = ADyp * (0.25 * clight * laser_norm)
        ux, uy, psi = pl.ux_half[sel].copy(), pl.uy_half[sel].copy(), pl.psi_half[sel].copy()

        def substep(ux, uy, psi):
            psi_inv = 1.0 / psi
            dux, duy, dps = _momentum_push(ux, uy, psi_inv, ExmByp, EypBxp, Ezp, Bxp, Byp, Bzp,
                                           clight_inv, qmc, Ap, ADxp, ADyp)
            duxe, duye, dpse = _momentum_push_dual(
                ux, dux, uy, duy, psi_inv, -psi_inv * psi_inv * dps,
                ExmByp, EypBxp, Ezp, Bxp, Byp, Bzp, clight_inv, qmc, Ap, ADxp, ADyp)
            ux = ux + (sdz * dux + 0.5 * sdz * sdz * duxe)
            uy = uy + (sdz * duy + 0.5 * sdz * sdz * duye)
            psi = psi + (sdz * dps + 0.5 * sdz * sdz * dpse)
            return ux, uy, psi
        for _i in range(nsub):                                                  # :148-168
            ux, uy, psi = substep(ux, uy, psi)
        xp = xp + dz * clight_inv * (ux * (1.0 / psi))                           # :173-174
        yp = yp + dz * clight_inv * (uy * (1.0 / psi))
        invalid = enforce_bc(xp, yp, ux, uy, bc_lo, bc_hi, bc_kind)             # :176
        if invalid.any():
            dead = sel[invalid]
            pl.w[dead] = 0.0
            pl.valid[dead] = False
            keep = ~invalid
            sel, xp, yp, ux, uy, psi = sel[keep], xp[keep], yp[keep], ux[keep], uy[keep], psi[keep]
            ExmByp, EypBxp, Ezp = ExmByp[keep], EypBxp[keep], Ezp[keep]
            Bxp, Byp, Bzp = Bxp[keep], Byp[keep], Bzp[keep]
            if aabs is not None:
                Ap, ADxp, ADyp = Ap[keep], ADxp[keep], ADyp[keep]
        pl.x[sel], pl.y[sel] = xp, yp
        if not temp_slice:
            pl.ux_half[sel], pl.uy_half[sel], pl.psi_half[sel] = ux, uy, psi
            pl.x_prev[sel], pl.y_prev[sel] = xp, yp
        for _i in range(nsub // 2):                                             # :194-214
            ux, uy, psi = substep(ux, uy, psi)
        pl.ux[sel], pl.uy[sel], pl.psi[sel] = ux, uy, psi


def advance_beam_slice(bs: dict, beam: Beam, F, geom: Geometry, pc: PhysConst, islice: int,
                       dt_step: float, time: float, bc_kind: str, bc_lo, bc_hi):
    """AdvanceBeamParticlesSlice, particles/pusher/BeamParticleAdvance.cpp:19-336 (lev 0, no
    radiation reaction, no spin).  Pushes every particle of the slice INCLUDING the slipped ones
    (:127), each from its own sub-cycle counter bs['nsub'] (:145, :333)."""
    if bs is None or bs['x'].size == 0:
        return
    n_sub = beam.n_subcycles
    dt = dt_step / n_sub                                                          # :32
    clight, inv_clight = pc.c, 1.0 / pc.c
    inv_c2 = 1.0 / (pc.c * pc.c)
    cmr = beam.charge / beam.mass                                                 # :99
    min_z = geom.lo[2] + islice * geom.dz                                         # :100
    xp, yp, zp = bs['x'].copy(), bs['y'].copy(), bs['z'].copy()
    ux, uy, uz = bs['ux'].copy(), bs['uy'].copy(), bs['uz'].copy()
    nsub = bs['nsub'].copy()
    alive = bs['valid'].copy()            # particles that still write back at the end
    for it in range(n_sub):
        act = alive & (nsub == it) & ~(zp < min_z)                                # :147-152
        if not act.any():
            continue
        a = np.nonzero(act)[0]
        x, y, z = xp[a], yp[a], zp[a]
        vx, vy, vz = ux[a], uy[a], uz[a]
        gammap_inv = 1.0 / np.sqrt(1.0 + (vx * vx + vy * vy + vz * vz) * inv_c2)  # :154-155
        x = x + dt * 0.5 * vx * gammap_inv                                        # :159-160
        y = y + dt * 0.5 * vy * gammap_inv
        inv = enforce_bc(x, y, vx, vy, bc_lo, bc_hi, bc_kind)                     # :162
        if inv.any():
            dead = a[inv]
            bs['w'][dead] = 0.0
            bs['valid'][dead] = False
            alive[dead] = False
            k = ~inv
            a, x, y, z, vx, vy, vz, gammap_inv = a[k], x[k], y[k], z[k], vx[k], vy[k], vz[k], gammap_inv[k]
        ExmByp, EypBxp, Ezp, Bxp, Byp, Bzp = gather_fields(x, y, F, geom)         # :192-194
        if beam.external_fields is not None:                                      # ExternalFields.H:29-58
            Ex, Ey, Ez_, Bx_, By_, Bz_ = beam.external_fields(x, y, z, time)
            ExmByp = ExmByp + (Ex - clight * By_)
            EypBxp = EypBxp + (Ey + clight * Bx_)
            Ezp = Ezp + Ez_
            Bxp = Bxp + Bx_
            Byp = Byp + By_
            Bzp = Bzp + Bz_
        ux_next = vx + dt * cmr * (ExmByp + (clight - vz * gammap_inv) * Byp + vy * gammap_inv * Bzp)
        uy_next = vy + dt * cmr * (EypBxp + (vz * gammap_inv - clight) * Bxp - vx * gammap_inv * Bzp)
        ux_i = (ux_next + vx) * 0.5                                               # :208-211
        uy_i = (uy_next + vy) * 0.5
        uz_i = vz + dt * 0.5 * cmr * Ezp
        gi_inv = 1.0 / np.sqrt(1.0 + (ux_i * ux_i + uy_i * uy_i + uz_i * uz_i) * inv_c2)
        uz_next = vz + dt * cmr * (Ezp + (ux_i * Byp - uy_i * Bxp) * gi_inv)      # :240-242
        gn_inv = 1.0 / np.sqrt(1.0 + (ux_next * ux_next + uy_next * uy_next + uz_next * uz_next) * inv_c2)
        x = x + dt * 0.5 * ux_next * gn_inv                                       # :312-314
        y = y + dt * 0.5 * uy_next * gn_inv
        if beam.do_z_push:
            z = z + dt * (uz_next * gn_inv - clight)
        xp[a], yp[a], zp[a] = x, y, z
        ux[a], uy[a], uz[a] = ux_next, uy_next, uz_next
        nsub[a] += 1
    a = np.nonzero(alive)[0]
    x, y, vx, vy = xp[a], yp[a], ux[a], uy[a]
    inv = enforce_bc(x, y, vx, vy, bc_lo, bc_hi, bc_kind)                         # :319
    if inv.any():
        dead = a[inv]
        bs['w'][dead] = 0.0
        bs['valid'][dead] = False
        k = ~inv
        a, x, y, vx, vy = a[k], x[k], y[k], vx[k], vy[k]
    bs['x'][a], bs['y'][a], bs['z'][a] = x, y, zp[a]
    bs['ux'][a], bs['uy'][a], bs['uz'][a] = vx, vy, uz[a]
    bs['nsub'][a] = nsub[a]


INSITU_NAMES = ('sum(w)', '[x]', '[x^2]', '[y]', '[y^2]', '[z]', '[z^2]', '[ux]', '[ux^2]', '[uy]', '[uy^2]',
                '[uz]', '[uz^2]', '[x*ux]', '[y*uy]', '[z*uz]', '[x*uy]', '[y*ux]', '[ux/uz]', '[uy/uz]',
                '[ga]', '[ga^2]')


def beam_insitu_sums(bs: dict, pc: PhysConst, radius=math.inf):
    """BeamParticleContainer::InSituComputeDiags, particles/beam/BeamParticleContainer.cpp:476-557:
    the 22 raw weighted sums + Np of one beam slice (getNumParticles: without the slipped ones)"""
    out = np.zeros(23)
    if bs is None or bs['x'].size == 0:
        return out
    n = bs.get('np', bs['x'].size)
    x, y, z, w = bs['x'][:n], bs['y'][:n], bs['z'][:n], bs['w'][:n]
    ci = 1.0 / pc.c
    ux, uy, uz = bs['ux'][:n] * ci, bs['uy'][:n] * ci, bs['uz'][:n] * ci
    sel = bs['valid'][:n] & ~(x * x + y * y > radius * radius)
    x, y, z, w, ux, uy, uz = (a[sel] for a in (x, y, z, w, ux, uy, uz))
    uz_inv = np.where(uz == 0.0, 0.0, 1.0 / np.where(uz == 0.0, 1.0, uz))
    ga = np.sqrt(1.0 + ux * ux + uy * uy + uz * uz)
    terms = (w, w * x, w * x * x, w * y, w * y * y, w * z, w * z * z, w * ux, w * ux * ux, w * uy,
             w * uy * uy, w * uz, w * uz * uz, w * x * ux, w * y * uy, w * z * uz, w * x * uy, w * y * ux,
             w * ux * uz_inv, w * uy * uz_inv, w * ga, w * ga * ga)
    for k, t in enumerate(terms):
        out[k] = t.sum()
    out[22] = x.size
    return out


PLASMA_INSITU_NAMES = ('sum(w)', '[x]', '[x^2]', '[y]', '[y^2]', '[ux]', '[ux^2]', '[uy]', '[uy^2]', '[uz]',
                       '[uz^2]', '[ga]', '[ga^2]', '[(ga-1)*(1-vz)]')


def plasma_insitu_sums(pl: 'Plasma', pc: PhysConst, radius=math.inf):
    """PlasmaParticleContainer::InSituComputeDiags, particles/plasma/PlasmaParticleContainer.cpp:443-526"""
    ci = 1.0 / pc.c
    sel = pl.valid & ~(pl.x * pl.x + pl.y * pl.y > radius * radius)
    x, y, psi, w0 = pl.x[sel], pl.y[sel], pl.psi[sel], pl.w[sel]
    ux, uy = pl.ux[sel] * ci, pl.uy[sel] * ci
    ga = (1.0 + ux * ux + uy * uy + psi * psi) / (2.0 * psi)
    uz = ga - psi
    w = w0 * ga / psi
    terms = (w, w * x, w * x * x, w * y, w * y * y, w * ux, w * ux * ux, w * uy, w * uy * uy, w * uz,
             w * uz * uz, w * ga, w * ga * ga, w0 * (ga - 1.0))
    out = np.zeros(15)
    for k, t in enumerate(terms):
        out[k] = t.sum()
    out[14] = x.size
    return out


def insitu_plasma_record(sums, time, step, charge, mass, z_lo, z_hi, density_factor, normalized):
    """PlasmaParticleContainer::InSituWriteToFile, PlasmaParticleContainer.cpp:530-618"""
    ns = sums.shape[1]
    sw = sums[0]
    sw_inv = np.where(sw <= 0.0, 0.0, 1.0 / np.where(sw <= 0.0, 1.0, sw))
    keep = (np.arange(14) == 0) | (np.arange(14) == 13)
    per = sums[:14] * np.where(keep[:, None], 1.0, sw_inv[None, :])
    tot = np.zeros(14)
    for isl in range(ns - 1, -1, -1):
        tot += sums[:14, isl]
    avg = tot / tot[0]
    N = PLASMA_INSITU_NAMES
    dt = np.dtype([('time', '<f8'), ('step', '<i4'), ('n_slices', '<i4'), ('charge', '<f8'), ('mass', '<f8'),
                   ('z_lo', '<f8'), ('z_hi', '<f8'), ('normalized_density_factor', '<f8'),
                   ('is_normalized_units', '<i4')] + [(nm, '<f8', (ns,)) for nm in N[1:]]
                  + [('sum(w)', '<f8', (ns,)), ('Np', '<i4', (ns,)),
                     ('average', [(nm, '<f8') for nm in N[1:13]]),
                     ('total', [('sum(w)', '<f8'), (N[13], '<f8'), ('Np', '<i4')])])
    rec = np.zeros((), dtype=dt)
    rec['time'], rec['step'], rec['n_slices'] = time, step, ns
    rec['charge'], rec['mass'], rec['z_lo'], rec['z_hi'] = charge, mass, z_lo, z_hi
    rec['normalized_density_factor'], rec['is_normalized_units'] = density_factor, int(normalized)
    for k, nm in enumerate(N):
        rec[nm] = per[k]
        if 1 <= k <= 12:
            rec['average'][nm] = avg[k]
    rec['Np'] = sums[14].astype(np.int32)
    rec['total']['sum(w)'] = tot[0]
    rec['total'][N[13]] = tot[13]
    rec['total']['Np'] = int(sums[14].sum())
    return dt, rec


FIELD_INSITU_NAMES = ('[Ex^2]', '[Ey^2]', '[Ez^2]', '[Bx^2]', '[By^2]', '[Bz^2]', '[ExmBy^2]', '[EypBx^2]',
                      '[jz_beam]', '[Ez*jz_beam]')


def field_insitu_sums(F, geom: 'Geometry', pc: PhysConst):
    """Fields::InSituComputeDiags, fields/Fields.cpp:1289-1347 (raw sums over the valid box)"""
    G = geom.g
    v = (slice(G, -G), slice(G, -G))
    T = lambda n: F[('This', n)][v]
    ex, ey = T('ExmBy') + T('By') * pc.c, T('EypBx') - T('Bx') * pc.c
    terms = (ex * ex, ey * ey, T('Ez') ** 2, T('Bx') ** 2, T('By') ** 2, T('Bz') ** 2, T('ExmBy') ** 2,
             T('EypBx') ** 2, T('jz_beam'), T('Ez') * T('jz_beam'))
    return np.array([t.sum() for t in terms])


def insitu_field_record(sums, time, step, z_lo, z_hi, normalized, dxdydz):
    """Fields::InSituWriteToFile, fields/Fields.cpp:1349-1428"""
    ns = sums.shape[1]
    per = sums * dxdydz
    tot = np.zeros(10)
    for isl in range(ns - 1, -1, -1):
        tot += per[:, isl]
    N = FIELD_INSITU_NAMES
    dt = np.dtype([('time', '<f8'), ('step', '<i4'), ('n_slices', '<i4'), ('z_lo', '<f8'), ('z_hi', '<f8'),
                   ('is_normalized_units', '<i4')] + [(nm, '<f8', (ns,)) for nm in N]
                  + [('integrated', [(nm, '<f8') for nm in N])])
    rec = np.zeros((), dtype=dt)
    rec['time'], rec['step'], rec['n_slices'] = time, step, ns
    rec['z_lo'], rec['z_hi'], rec['is_normalized_units'] = z_lo, z_hi, int(normalized)
    for k, nm in enumerate(N):
        rec[nm] = per[k]
        rec['integrated'][nm] = tot[k]
    return dt, rec


LASER_INSITU_NAMES = ('max(|a|^2)', '[|a|^2]', '[|a|^2*x]', '[|a|^2*x*x]', '[|a|^2*y]', '[|a|^2*y*y]')


def laser_insitu_sums(env, geom: 'Geometry'):
    """MultiLaser::InSituComputeDiags, laser/MultiLaser.cpp:923-1001: raw values of one slice --
    max |a|^2, five sums, and the sum of a over the centre cell(s) (re, im)"""
    ny, nx = env.shape
    a2 = env.real ** 2 + env.imag ** 2
    x = (np.arange(nx) * geom.dx + geom.pos_offset(0))[None, :]
    y = (np.arange(ny) * geom.dy + geom.pos_offset(1))[:, None]
    xs = sorted({(nx - 1) // 2, nx // 2})
    ys = sorted({(ny - 1) // 2, ny // 2})
    axis = env[np.ix_(ys, xs)].sum()
    return np.array([a2.max(), a2.sum(), (a2 * x).sum(), (a2 * x * x).sum(), (a2 * y).sum(),
                     (a2 * y * y).sum(), axis.real, axis.imag])


def insitu_laser_record(sums, time, step, z_lo, z_hi, normalized, dxdydz, nx, ny):
    """MultiLaser::InSituWriteToFile, laser/MultiLaser.cpp:1003-1075"""
    ns = sums.shape[1]
    mid = (1.0 if (nx - 1) // 2 == nx // 2 else 0.5) * (1.0 if (ny - 1) // 2 == ny // 2 else 0.5)
    per = sums[:6].copy()
    per[1:] *= dxdydz
    tot = np.zeros(6)
    for isl in range(ns - 1, -1, -1):
        tot[0] = max(tot[0], per[0, isl])
        tot[1:] += per[1:, isl]
    N = LASER_INSITU_NAMES
    dt = np.dtype([('time', '<f8'), ('step', '<i4'), ('n_slices', '<i4'), ('z_lo', '<f8'), ('z_hi', '<f8'),
                   ('is_normalized_units', '<i4')] + [(nm, '<f8', (ns,)) for nm in N]
                  + [('axis(a)', '<c16', (ns,)), ('integrated', [(nm, '<f8') for nm in N])])
    rec = np.zeros((), dtype=dt)
    rec['time'], rec['step'], rec['n_slices'] = time, step, ns
    rec['z_lo'], rec['z_hi'], rec['is_normalized_units'] = z_lo, z_hi, int(normalized)
    for k, nm in enumerate(N):
        rec[nm] = per[k]
        rec['integrated'][nm] = tot[k]
    rec['axis(a)'] = (sums[6] + 1j * sums[7]) * mid
    return dt, rec


def insitu_beam_record(sums, time, step, charge, mass, z_lo, z_hi, density_factor, normalized):
    """InSituWriteToFile, particles/beam/BeamParticleContainer.cpp:596-732: (numpy dtype, record)
    of one time step from the raw sums[23, n_slices]; format of utils/InsituUtil.H"""
    ns = sums.shape[1]
    sw = sums[0]
    sw_inv = np.where(sw <= 0.0, 0.0, 1.0 / np.where(sw <= 0.0, 1.0, sw))
    per = sums[:22] * np.where(np.arange(22)[:, None] == 0, 1.0, sw_inv[None, :])
    tot = np.zeros(22)
    for isl in range(ns - 1, -1, -1):      # accumulated slice by slice from the head (:548)
        tot += sums[:22, isl]
    avg = tot / tot[0]
    arr = [(nm, '<f8', (ns,)) for nm in INSITU_NAMES[1:]] + [('sum(w)', '<f8', (ns,)), ('Np', '<i4', (ns,))]
    dt = np.dtype([('time', '<f8'), ('step', '<i4'), ('n_slices', '<i4'), ('charge', '<f8'), ('mass', '<f8'),
                   ('z_lo', '<f8'), ('z_hi', '<f8'), ('normalized_density_factor', '<f8'),
                   ('is_normalized_units', '<i4')] + arr
                  + [('average', [(nm, '<f8') for nm in INSITU_NAMES[1:]]),
                     ('total', [('sum(w)', '<f8'), ('Np', '<i4')])])
    rec = np.zeros((), dtype=dt)
    rec['time'], rec['step'], rec['n_slices'] = time, step, ns
    rec['charge'], rec['mass'], rec['z_lo'], rec['z_hi'] = charge, mass, z_lo, z_hi
    rec['normalized_density_factor'], rec['is_normalized_units'] = density_factor, int(normalized)
    for k, nm in enumerate(INSITU_NAMES):
        rec[nm] = per[k]
        if k:
            rec['average'][nm] = avg[k]
    rec['Np'] = sums[22].astype(np.int32)
    rec['total']['sum(w)'] = tot[0]
    rec['total']['Np'] = int(sums[22].sum())
    return dt, rec


_BEAM_KEYS = ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz', 'id', 'valid', 'nsub')


def empty_beam_slice():
    bs = {k: np.zeros(0) for k in ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz')}
    bs['id'] = np.zeros(0, dtype=np.int64)
    bs['valid'] = np.zeros(0, dtype=bool)
    bs['nsub'] = np.zeros(0, dtype=np.int64)
    bs['np'] = 0
    return bs


def shift_slipped_particles(bs_this: dict, bs_next: dict | None, geom: Geometry, islice: int):
    """shiftSlippedParticles, particles/sorting/SliceSort.cpp:13-67: drop invalid particles,
    keep those with z >= min_z (stable order), append the others to the Next slice as 'slipped'
    (they are pushed there with their remaining sub-cycles but not deposited, see
    BeamParticleContainer.H:175-181).  PARITY UNPINNED (no deterministic golden slips)."""
    if bs_this is None or bs_this['x'].size == 0:
        return
    min_z = geom.lo[2] + islice * geom.dz
    v = bs_this['valid']
    stay = v & (bs_this['z'] >= min_z)
    slip = v & ~stay
    if bs_next is not None and slip.any():
        assert bs_next['x'].size == bs_next['np']          # SliceSort.cpp:45
        for k in _BEAM_KEYS:
            bs_next[k] = np.concatenate([bs_next[k], bs_this[k][slip]])
    for k in _BEAM_KEYS:
        bs_this[k] = bs_this[k][stay]
    bs_this['np'] = int(stay.sum())


# --------------------------------------------------------------------------------------------
# Initialisation
# --------------------------------------------------------------------------------------------

def init_plasma(pl: Plasma, geom: Geometry, pc: PhysConst, normalized: bool, bc_lo, bc_hi,
                c_t: float = 0.0):
    """PlasmaParticleContainer::InitParticles, particles/plasma/PlasmaParticleContainerInit.cpp:
    17-316 (GPU order: ppc index outermost, cells x-fastest; no fine patch, u_std = 0)."""
    dx, dy, dz = geom.dx, geom.dy, geom.dz
    ppcx, ppcy = pl.ppc
    nppc = ppcx * ppcy
    scale = 0.0 if nppc <= 0 else (1.0 / nppc if normalized else dx * dy * dz / nppc)   # :40-41
    ilo, ihi, jlo, jhi = 0, geom.nx - 1, 0, geom.ny - 1
    if pl.radius != math.inf:                                                  # :70-82
        ilo = max(ilo, int(round((-pl.radius - geom.lo[0]) / dx - 2)))
        jlo = max(jlo, int(round((-pl.radius - geom.lo[1]) / dy - 2)))
        ihi = min(ihi, int(round((pl.radius - geom.lo[0]) / dx + 2)))
        jhi = min(jhi, int(round((pl.radius - geom.lo[1]) / dy + 2)))
    jj, ii = np.meshgrid(np.arange(jlo, jhi + 1), np.arange(ilo, ihi + 1), indexing='ij')
    ii, jj = ii.ravel().astype(float), jj.ravel().astype(float)
    xs, ys, ws = [], [], []
    for i_part in range(nppc):
        rx = (0.5 + (i_part % ppcx)) / ppcx                                    # ParticleUtil.H:72-80
        ry = (0.5 + (i_part // ppcx)) / ppcy
        x = geom.lo[0] + (ii + rx) * dx
        y = geom.lo[1] + (jj + ry) * dy
        rsq = x * x + y * y
        dens = np.broadcast_to(np.asarray(pl.density(x, y, c_t), dtype=float), x.shape)
        keep = ~((x >= bc_hi[0]) | (x < bc_lo[0]) | (y >= bc_hi[1]) | (y < bc_lo[1])
                 | (rsq > pl.radius * pl.radius)
                 | (rsq < pl.hollow_core_radius * pl.hollow_core_radius)
                 | (dens <= pl.min_density))                                     # :162-166
        xs.append(x[keep]); ys.append(y[keep]); ws.append(dens[keep] * scale)
    pl.x = np.concatenate(xs); pl.y = np.concatenate(ys); pl.w = np.concatenate(ws)
    n = pl.x.size
    u = pl.u_mean
    pl.ux = np.full(n, u[0] * pc.c); pl.uy = np.full(n, u[1] * pc.c)
    pl.psi = np.full(n, math.sqrt(1.0 + u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) - u[2])
    pl.x_prev, pl.y_prev = pl.x.copy(), pl.y.copy()
    pl.ux_half, pl.uy_half, pl.psi_half = pl.ux.copy(), pl.uy.copy(), pl.psi.copy()
    pl.valid = np.ones(n, dtype=bool)


def beam_density(beam: Beam, x, y, z):
    """GetInitialDensity, particles/profiles/GetInitialDensity.H:33-51"""
    if beam.profile == 'gaussian':
        dxn = (x - beam.position_mean[0]) / beam.position_std[0]
        dyn = (y - beam.position_mean[1]) / beam.position_std[1]
        dzn = (z - beam.position_mean[2]) / beam.position_std[2]
        return beam.density * np.exp(-0.5 * dxn * dxn) * np.exp(-0.5 * dyn * dyn) \
            * np.exp(-0.5 * dzn * dzn)
    if beam.profile == 'flattop':
        return np.full_like(x, beam.density)
    raise NotImplementedError(beam.profile)


def init_beam_slice(beam: Beam, islice: int, geom: Geometry, pc: PhysConst, normalized: bool):
    """BeamParticleContainer::InitBeamFixedPPCSlice,
    particles/beam/BeamParticleContainerInit.cpp:198-346 (random_ppc = 0 0 0, u_std = 0)."""
    dx, dy, dz = geom.dx, geom.dy, geom.dz
    px, py, pz = beam.ppc
    nppc = px * py * pz
    scale = 1.0 / nppc if normalized else dx * dy * dz / nppc
    jj, ii = np.meshgrid(np.arange(geom.ny), np.arange(geom.nx), indexing='ij')
    ii, jj = ii.ravel().astype(float), jj.ravel().astype(float)
    cols = []
    for i_part in range(nppc):
        ix_p = i_part // (py * pz)                                   # ParticleUtil.H:49-63
        iy_p = (i_part % (py * pz)) % py
        iz_p = (i_part % (py * pz)) // py
        x = geom.lo[0] + (ii + (0.5 + ix_p) / px) * dx
        y = geom.lo[1] + (jj + (0.5 + iy_p) / py) * dy
        z = np.full_like(x, geom.lo[2] + (islice + (0.5 + iz_p) / pz) * dz)
        xm, ym = beam.position_mean[0], beam.position_mean[1]
        keep = ~((z >= beam.zmax) | (z < beam.zmin)
                 | (((x - xm) * (x - xm) + (y - ym) * (y - ym)) > beam.radius * beam.radius))
        dens = beam_density(beam, x, y, z)
        keep &= ~(dens <= beam.min_density)
        cols.append((keep, x, y, z, dens * scale))
    # particle order: cell-major (x fastest), i_part inner (:298-345)
    keep = np.stack([c[0] for c in cols], axis=1).ravel()
    pick = lambda k: np.stack([c[k] for c in cols], axis=1).ravel()[keep]
    n = int(keep.sum())
    bs = dict(x=pick(1), y=pick(2), z=pick(3), w=np.abs(pick(4)),
              ux=np.full(n, beam.u_mean[0] * pc.c), uy=np.full(n, beam.u_mean[1] * pc.c),
              uz=np.full(n, beam.u_mean[2] * pc.c),
              id=np.arange(beam.next_id, beam.next_id + n, dtype=np.int64),
              valid=np.ones(n, dtype=bool), nsub=np.zeros(n, dtype=np.int64))
    bs['np'] = n                 # particles without the slipped ones (BeamParticleContainer.H:175)
    beam.next_id += n
    return bs


# --------------------------------------------------------------------------------------------
# Driver (Hipace.cpp:393-728)
# --------------------------------------------------------------------------------------------

class Simulation:
    def __init__(self, deck_text: str, overrides: dict | None = None, numprocs: int = 1):
        """numprocs: the number of pipeline ranks whose adaptive-time-step bookkeeping is emulated (rank
        r owns the steps r, r + numprocs, ...: each keeps its OWN dt and min_uz_mq, the physical time
        travels with the beam, MultiBuffer::get_time / put_time, utils/MultiBuffer.cpp:611-651).  Everything
        else of a run is independent of the number of ranks."""
        self.numprocs = int(numprocs)
        d = self.deck = parse_deck(deck_text, overrides)
        self.normalized = bool(_get(d, 'hipace.normalized_units', 0, typ=int))
        self.pc = PhysConst.make(self.normalized)
        n = _get(d, 'amr.n_cell', n=3, typ=int)
        lo = _get(d, 'geometry.prob_lo', n=3)
        hi = _get(d, 'geometry.prob_hi', n=3)
        order = _get(d, 'hipace.depos_order_xy', 2, typ=int)
        dtype = _get(d, 'hipace.depos_derivative_type', 2, typ=int)
        assert 0 <= order <= 3 and 0 <= dtype <= 2
        assert order != 0 or dtype != 0, 'Analytic derivative with depos_order=0 would vanish'  # Hipace.cpp:52-53
        self.geom = Geometry(n[0], n[1], n[2], tuple(lo), tuple(hi), order, dtype)
        self.diag_type = _get(d, 'diagnostic.diag_type', 'xyz', typ=str)
        assert self.diag_type in ('xyz', 'xz'), 'oracle scope: xyz / xz field diagnostics'
        solver = _get(d, 'hipace.bxby_solver', 'explicit', typ=str)
        assert solver in ('explicit', 'predictor-corrector')
        self.explicit = solver == 'explicit'
        self.predcorr_tol = _get(d, 'hipace.predcorr_B_error_tolerance', 4e-2)      # Hipace.H:210-222
        self.predcorr_max_iter = _get(d, 'hipace.predcorr_max_iterations', 30, typ=int)
        self.predcorr_mix = _get(d, 'hipace.predcorr_B_mixing_factor', 0.05)
        self.predcorr_iters = []
        bf = _get(d, 'boundary.field', typ=str)
        assert bf in ('Dirichlet', 'Open', 'Periodic')
        assert bf != 'Open' or not self.explicit, 'oracle scope: Open only with predictor-corrector'
        self.open_bc = bf == 'Open'
        self.field_periodic = bf == 'Periodic'
        ps = _get(d, 'fields.poisson_solver', 'FFTDirichletFast', typ=str)     # fields/Fields.cpp:34-40, :179-208
        assert ps in ('FFTDirichletFast', 'FFTDirichletDirect', 'FFTDirichletExpanded', 'FFTPeriodic')
        self.poisson_periodic = ps == 'FFTPeriodic'
        assert self.explicit or not (self.field_periodic or self.poisson_periodic), \
            'oracle scope: periodic fields only with the explicit solver'
        self.bc_kind = _get(d, 'boundary.particle', typ=str)
        self.bc_lo = _get(d, 'boundary.particle_lo', [lo[0], lo[1]], n=2)
        self.bc_hi = _get(d, 'boundary.particle_hi', [hi[0], hi[1]], n=2)
        self.max_step = _get(d, 'max_step', 0, typ=int)
        # hipace.dt = adaptive (utils/AdaptiveTimeStep.cpp:17-60; defaults AdaptiveTimeStep.H:25-51)
        self.adaptive_dt = str(d.get('hipace.dt', ['0'])[0]) == 'adaptive'
        self.dt = 0.0 if self.adaptive_dt else _get(d, 'hipace.dt', 0.0)
        self.nt_per_betatron = _get(d, 'hipace.nt_per_betatron', 20.0)
        self.dt_max = _get(d, 'hipace.dt_max', math.inf)
        self.adaptive_threshold_uz = _get(d, 'hipace.adaptive_threshold_uz', 2.0)
        self.adaptive_density = _get(d, 'plasmas.adaptive_density', 0.0)
        self.adaptive_phase_tolerance = _get(d, 'hipace.adaptive_phase_tolerance', 4e-4)
        self.adaptive_phase_substeps = _get(d, 'hipace.adaptive_phase_substeps', 2000, typ=int)
        self.adaptive_control_phase = bool(_get(d, 'hipace.adaptive_control_phase_advance', 1, typ=int))
        self.adaptive_predict_step = bool(_get(d, 'hipace.adaptive_predict_step', 1, typ=int))
        assert not (self.adaptive_dt and _get(d, 'hipace.adaptive_gather_ez', 0, typ=int)), \
            'oracle scope: hipace.adaptive_gather_ez = 0' 
        self.mg_tol_rel = _get(d, 'hipace.MG_tolerance_rel', 1e-4)
        self.mg_tol_abs = _get(d, 'hipace.MG_tolerance_abs', np.finfo(float).tiny)
        fd = d.get('diagnostic.field_data', [])
        self.deposit_rho = ('rho' in fd) or bool(_get(d, 'hipace.deposit_rho', 0, typ=int))
        self.do_beam_jx_jy = bool(_get(d, 'hipace.do_beam_jx_jy_deposition', 1, typ=int))

        self.plasmas = []
        names = d.get('plasmas.names', ['no_plasma'])
        if names[0] != 'no_plasma':
            for nm in names:
                self.plasmas.append(self._read_plasma(nm))
        self.beams = []
        bnames = d.get('beams.names', ['no_beam'])
        if bnames[0] != 'no_beam':
            for nm in bnames:
                self.beams.append(self._read_beam(nm))
        self.any_neutral = any(p.neutralize_background for p in self.plasmas)
        self.grid_current = None                                  # utils/GridCurrent.cpp:14-23
        if _get(d, 'grid_current.use_grid_current', 0, typ=int):
            self.grid_current = (_get(d, 'grid_current.peak_current_density'),
                                 _get(d, 'grid_current.position_mean', n=3),
                                 _get(d, 'grid_current.position_std', n=3))
        # lasers (laser/MultiLaser.cpp:26-56, laser/Laser.cpp:18-47): gaussian envelopes on the
        # field grid; only what time step 0 needs (no envelope advance)
        self.lasers = []
        lnames = d.get('lasers.names', ['no_laser'])
        self.use_laser = lnames[0] != 'no_laser'
        if self.use_laser:
            self.laser_lambda0 = _get(d, 'lasers.lambda0')
            self.laser_interp_order = _get(d, 'lasers.interp_order', 1, typ=int)
            for k in ('lasers.n_cell', 'lasers.patch_lo', 'lasers.patch_hi'):
                if k in d:
                    raise NotImplementedError('oracle scope: laser grid = field grid (' + k + ')')
            for nm in lnames:
                if _get(d, nm + '.init_type', 'gaussian', typ=str) != 'gaussian':
                    raise NotImplementedError('oracle scope: gaussian lasers')
                has_L0, has_tau = (nm + '.L0') in d, (nm + '.tau') in d
                assert has_L0 != has_tau, 'specify exclusively L0 or tau'          # Laser.cpp:38-41
                L0 = _get(d, nm + '.L0') if has_L0 else _get(d, nm + '.tau') * self.pc.c
                self.lasers.append(Laser(
                    name=nm, a0=_get(d, nm + '.a0', 0.0), w0=_get(d, nm + '.w0', 0.0),
                    cep=_get(d, nm + '.CEP', 0.0),
                    propagation_angle_yz=_get(d, nm + '.propagation_angle_yz', 0.0),
                    pft_yz=_get(d, nm + '.PFT_yz', math.pi / 2.0), L0=L0,
                    focal_distance=_get(d, nm + '.focal_distance', 0.0),
                    position_mean=tuple(_get(d, nm + '.position_mean', [0., 0., 0.], n=3))))
            self.laser_solver = _get(d, 'lasers.solver_type', 'multigrid', typ=str)
            self.laser_use_phase = bool(_get(d, 'lasers.use_phase', 1, typ=int))
            assert self.laser_solver in ('fft', 'multigrid')
            self.laser_mg_tol_rel = _get(d, 'lasers.MG_tolerance_rel', 1e-4)
            self.laser_mg_tol_abs = _get(d, 'lasers.MG_tolerance_abs', 0.0)
            self.laser_mg_avg_rhs = bool(_get(d, 'lasers.MG_average_rhs', 1, typ=int))
            self.laser_mg = None
            self.laser_store = {}           # islice -> (A^n, A^{n-1}) handed from step to step
            self.laser_next = {}
        self.comps, self.ncomp = component_map(self.deposit_rho, self.any_neutral, self.use_laser,
                                               self.explicit)
        g = self.geom
        self.F = {k: np.zeros((g.ny + 2 * g.g, g.nx + 2 * g.g)) for k in self.comps}
        self.eig = poisson_eigenvalues(g.nx, g.ny, g.dx, g.dy)
        self.mg = MultiGrid1(g.dx, g.dy, g.nx, g.ny)
        self.checksums = {}
        self.n_plasma_pushed = 0
        self.n_beam_pushed = 0
        self.step = 0
        self.time = 0.0
        self.beam_diag = {}
        self.mg_cycles = []
        self.n_qsa_violation = 0
        self.slice_hook = None        # callable(sim, islice, stage) for tests
        # in-situ beam diagnostics ("<beam name> or beams", BeamParticleContainer.cpp:61-63)
        self.insitu_period = {b.name: _get(d, b.name + '.insitu_period',
                                           _get(d, 'beams.insitu_period', 0, typ=int), typ=int)
                              for b in self.beams}
        self.insitu_radius = {b.name: _get(d, b.name + '.insitu_radius',
                                           _get(d, 'beams.insitu_radius', math.inf))
                              for b in self.beams}
        self.insitu, self.insitu_records = {}, {}
        self.plasma_insitu_period = {p.name: _get(d, p.name + '.insitu_period',
                                                  _get(d, 'plasmas.insitu_period', 0, typ=int), typ=int)
                                     for p in self.plasmas}
        self.plasma_insitu_radius = {p.name: _get(d, p.name + '.insitu_radius',
                                                  _get(d, 'plasmas.insitu_radius', math.inf))
                                     for p in self.plasmas}
        self.plasma_insitu, self.plasma_insitu_records = {}, {}
        self.field_insitu_period = _get(d, 'fields.insitu_period', 0, typ=int)
        assert not (self.field_insitu_period and not self.explicit), \
            'Must use explicit solver for field insitu diagnostic'
        self.field_insitu, self.field_insitu_records = None, []
        self.laser_insitu_period = _get(d, 'lasers.insitu_period', 0, typ=int) if self.use_laser else 0
        self.laser_insitu, self.laser_insitu_records = None, []

    # -- deck readers ----------------------------------------------------------------------
    def _species_charge_mass(self, pre, default_element=None):
        d, pc = self.deck, self.pc
        el = _get(d, pre + '.element', default_element, typ=str)
        charge = mass = 0.0
        if el == 'electron':
            charge, mass = -pc.q_e, pc.m_e
        elif el == 'positron':
            charge, mass = pc.q_e, pc.m_e
        elif el == 'proton':
            charge, mass = pc.q_e, pc.m_p
        elif el:
            raise NotImplementedError('element ' + el)
        charge = _get(d, pre + '.charge', charge)
        mass = _get(d, pre + '.mass', mass)
        return charge, mass

    def _read_plasma(self, nm):
        """PlasmaParticleContainer::ReadParameters, particles/plasma/PlasmaParticleContainer.cpp:27-170"""
        d = self.deck
        charge, mass = self._species_charge_mass(nm)
        toks = _get(d, nm + '.density(x,y,z)', None, n=-1, typ=str, alt='plasmas.density(x,y,z)')
        expr = ' '.join(toks) if toks else '0.'        # the expression may contain blanks
        code = compile(expr.replace('^', '**'), '<density>', 'eval')
        env = _constants(d)
        env.update({f: getattr(np, f) for f in ('sqrt', 'exp', 'sin', 'cos', 'log', 'tanh')})

        def density(x, y, z, _code=code, _env=env):
            return eval(_code, {'__builtins__': {}}, dict(_env, x=x, y=y, z=z))
        ppc = _get(d, nm + '.ppc', n=2, typ=int, alt='plasmas.ppc')
        return Plasma(
            name=nm, charge=charge, mass=mass, ppc=tuple(ppc), density=density,
            neutralize_background=bool(_get(d, nm + '.neutralize_background', 1, typ=int,
                                            alt='plasmas.neutralize_background')),
            max_qsa_weighting_factor=_get(d, nm + '.max_qsa_weighting_factor', 35.,
                                          alt='plasmas.max_qsa_weighting_factor'),
            n_subcycles=_get(d, nm + '.n_subcycles', 1, typ=int, alt='plasmas.n_subcycles'),
            radius=_get(d, nm + '.radius', math.inf, alt='plasmas.radius'),
            hollow_core_radius=_get(d, nm + '.hollow_core_radius', 0., alt='plasmas.hollow_core_radius'),
            min_density=_get(d, nm + '.min_density', 0., alt='plasmas.min_density'),
            u_mean=tuple(_get(d, nm + '.u_mean', [0., 0., 0.], n=3)))

    def _read_beam(self, nm):
        """BeamParticleContainer::ReadParameters/InitData (fixed_ppc),
        particles/beam/BeamParticleContainer.cpp:41-143"""
        d = self.deck
        assert _get(d, nm + '.injection_type', typ=str) == 'fixed_ppc', 'oracle scope: fixed_ppc'
        charge, mass = self._species_charge_mass(nm, 'electron')
        return Beam(
            name=nm, charge=charge, mass=mass, ppc=tuple(_get(d, nm + '.ppc', [1, 1, 1], n=3, typ=int)),
            profile=_get(d, nm + '.profile', typ=str), density=abs(_get(d, nm + '.density')),
            zmin=_get(d, nm + '.zmin'), zmax=_get(d, nm + '.zmax'), radius=_get(d, nm + '.radius'),
            position_mean=tuple(_get(d, nm + '.position_mean', [0., 0., 0.], n=3)),
            position_std=tuple(_get(d, nm + '.position_std', [0., 0., 0.], n=3)),
            u_mean=tuple(_get(d, nm + '.u_mean', [0., 0., 0.], n=3)),
            min_density=abs(_get(d, nm + '.min_density', 0.)),
            n_subcycles=_get(d, nm + '.n_subcycles', 10, typ=int),
            do_z_push=bool(_get(d, nm + '.do_z_push', 1, typ=int)),
            external_fields=self._read_external_fields(nm))

    def _read_external_fields(self, nm):
        """external_E(x,y,z,t) / external_B(x,y,z,t), three expressions each,
        particles/beam/BeamParticleContainer.cpp:71-88 (beam name first, then 'beams')"""
        d = self.deck
        exprs = []
        used = False
        for key in ('external_E(x,y,z,t)', 'external_B(x,y,z,t)'):
            v = d.get(nm + '.' + key, d.get('beams.' + key))
            if v is None:
                exprs += ['0.', '0.', '0.']
            else:
                assert len(v) == 3, key + ' needs 3 expressions'
                exprs += list(v)
                used = True
        if not used:
            return None
        env = dict(_CONST_SI)
        env.update({f: getattr(np, f) for f in ('sqrt', 'exp', 'sin', 'cos', 'log', 'tanh')})
        for k, v in d.items():
            if k.startswith('my_constants.'):
                env[k.split('.', 1)[1]] = _eval(v[0], {kk: vv for kk, vv in d.items() if kk != k})
        codes = [compile(e.replace('^', '**'), '<external field>', 'eval') for e in exprs]

        def fields(x, y, z, t, _codes=codes, _env=env):
            loc = dict(_env, x=x, y=y, z=z, t=t)
            return tuple(np.broadcast_to(np.asarray(eval(c, {'__builtins__': {}}, loc), dtype=float),
                                         x.shape) for c in _codes)
        return fields

    # -- slice loop ------------------------------------------------------------------------
    def T(self, name):
        return self.F[('This', name)]

    def begin_step(self, step: int = 0):
        """Hipace::Evolve up to the slice loop, Hipace.cpp:401-475"""
        self.step = step
        if self.adaptive_dt:                                      # :411, :420, :434
            if step == 0 or not hasattr(self, '_rank_state'):
                # Hipace.cpp:275-281: the head rank's initial dt and min_uz_mq are broadcast to every rank
                # (a function of the deck only: a process that emulates one rank computes it itself)
                self.dt = 0.0
                self._adaptive_init()
                self._rank_state = {r: (self.dt, self._min_uz_mq) for r in range(self.numprocs)}
            # the rank that owns this step continues from ITS previous step's dt / min_uz_mq ...
            self.dt, self._min_uz_mq = self._rank_state[step % self.numprocs]
            # ... at the time its upstream neighbour computed (get_time :411, put_time :446)
            self.time = 0.0 if step == 0 else self._next_time
            self._adaptive_from_density()
            self._next_time = self.time + self.dt
        else:
            self.time = step * self.dt                            # :410, :434 (fixed dt)
        self.beam_diag = {b.name: [] for b in self.beams}
        for a in self.F.values():
            a[...] = 0.0                                          # ResetAllQuantities :730-742
        for b in self.beams:                                      # MultiBuffer: nsubcycles is not
            for bs in b.slices.values():                          # communicated (BeamParticleContainer.H:35-37)
                bs['nsub'][...] = 0
        for pl in self.plasmas:
            init_plasma(pl, self.geom, self.pc, self.normalized, self.bc_lo, self.bc_hi,
                        c_t=self.pc.c * self.time)
        if self.use_laser:                                        # :733-734, :452 SetInitialChi
            g = self.geom
            self.L = LaserSlices(g.ny, g.nx)
            x = (np.arange(g.nx) * g.dx + g.pos_offset(0))[None, :] + np.zeros((g.ny, 1))
            y = (np.arange(g.ny) * g.dy + g.pos_offset(1))[:, None] + np.zeros((1, g.nx))
            self.laser_chi_initial = np.zeros((g.ny, g.nx))
            for pl in self.plasmas:                               # laser/MultiLaser.cpp:293-332
                self.laser_chi_initial += pl.density(x, y, self.pc.c * self.time) \
                    * (pl.charge * pl.charge * self.pc.mu0 / pl.mass)
            self.laser_next = {}
        for pl in self.plasmas:                                   # :468-470, MultiPlasma.cpp:106-118
            if pl.neutralize_background:
                deposit_current(pl, self.F, self.geom, self.pc, self.normalized,
                                rhomjz=self.F[('RhomJzIons', 'rhomjz')], flip_charge=True)

    # ---- adaptive time step (utils/AdaptiveTimeStep.cpp), one rank ------------------------------
    def _max_charge_density(self, z):
        """MultiPlasma::maxChargeDensity, particles/plasma/MultiPlasma.cpp:63-73"""
        m = abs(self.adaptive_density * self.pc.q_e)
        for pl in self.plasmas:
            m = max(m, abs(pl.charge * float(pl.density(0.0, 0.0, z))))
        return m

    def _adaptive_init(self):
        """Hipace.cpp:275-279: the beam as specified in the deck (GatherMinUzSlice(initial), :96-106)"""
        self._ts = {}
        for b in self.beams:
            u, us = b.u_mean[2], 0.0                               # u_std = 0 in the oracle's scope
            self._ts[b.name] = dict(min_uz=u - 4.0 * us, sw=1.0, swu=u, swu2=u * u + us * us)
        self._min_uz_mq = np.finfo(float).max
        self._adaptive_from_min_uz(0.0)
        self._adaptive_from_density(at=0.0)

    def _adaptive_gather(self, islice):
        """GatherMinUzSlice(false) after the push of a slice, :108-141"""
        ci = 1.0 / self.pc.c
        for b in self.beams:
            bs = self.beam_slice(b, islice)
            n = bs['np']
            v = bs['valid'][:n]
            w, uz = bs['w'][:n][v], bs['uz'][:n][v]
            t = self._ts[b.name]
            t['sw'] += w.sum()
            t['swu'] += (w * uz * ci).sum()
            t['swu2'] += (w * uz * uz * ci * ci).sum()
            if uz.size:
                t['min_uz'] = min(t['min_uz'], float((uz * ci).min()))

    def _adaptive_from_min_uz(self, t_now):
        """CalculateFromMinUz, :143-233.  The new dt is used numprocs steps later: with
        hipace.adaptive_predict_step (default on) the betatron frequency is re-evaluated numprocs times
        at the predicted times (:233-251)"""
        niter = self.numprocs if self.adaptive_predict_step else 1
        new_dts, mq = [], []
        for b in self.beams:
            new_dt = self.dt
            t = self._ts[b.name]
            if b.charge == 0.0:
                new_dts.append(new_dt); mq.append(np.finfo(float).max)
                continue
            mcr = b.mass / b.charge
            assert t['sw'] != 0, 'The sum of all weights is 0'
            mean = t['swu'] / t['sw']
            sigma = math.sqrt(abs(t['swu2'] / t['sw'] - mean * mean))
            chosen = min(max(mean - 4.0 * sigma, t['min_uz']), 1e30)
            chosen = max(chosen, self.adaptive_threshold_uz)
            mq.append(abs(chosen * mcr))
            cand, new_time, min_uz = self.dt, t_now, chosen
            for _ in range(niter):
                rho = self._max_charge_density(self.pc.c * new_time)
                assert rho > 0.0, 'A >0 plasma density must be specified to use an adaptive time step.'
                min_uz = max(min_uz, 0.001 * self.adaptive_threshold_uz)
                omega_b = math.sqrt(rho / (2.0 * abs(min_uz * mcr) * self.pc.ep0))
                cand = 2.0 * math.pi / omega_b / self.nt_per_betatron
                new_time += cand
                if min_uz > self.adaptive_threshold_uz:
                    new_dt = cand
            new_dts.append(new_dt)
        self._min_uz_mq = min(mq)
        self.dt = min(min(new_dts), self.dt_max)

    def _adaptive_from_density(self, at=None):
        """CalculateFromDensity, :315-369: reset the gathered data; shorten dt if the betatron phase
        advance over the step deviates from that of a uniform plasma"""
        t0 = self.time if at is None else at
        for t in self._ts.values():
            t.update(min_uz=1e30, sw=0.0, swu=0.0, swu2=0.0)
        if not self.adaptive_control_phase:
            return
        dt_sub = self.dt / self.adaptive_phase_substeps
        pa = pa0 = 0.0
        omgb0 = math.sqrt(self._max_charge_density(self.pc.c * t0) / (2.0 * self._min_uz_mq * self.pc.ep0))
        for i in range(self.adaptive_phase_substeps):
            omgb = math.sqrt(self._max_charge_density(self.pc.c * (t0 + i * dt_sub))
                             / (2.0 * self._min_uz_mq * self.pc.ep0))
            pa += omgb * dt_sub
            pa0 += omgb0 * dt_sub
            if abs(pa - pa0) > 2.0 * math.pi * self.adaptive_phase_tolerance / self.nt_per_betatron:
                self.dt = i * dt_sub
                return

    def beam_slice(self, beam: Beam, islice: int):
        if islice < 0:
            return None
        if islice not in beam.slices:
            beam.slices[islice] = init_beam_slice(beam, islice, self.geom, self.pc, self.normalized)
        return beam.slices[islice]

    def _solve_one_slice_pc(self, islice: int):
        """Hipace::SolveOneSlice with hipace.bxby_solver = predictor-corrector: Hipace.cpp:556-728
        (the non-explicit branches) and PredictorCorrectorLoopToSolveBxBy, Hipace.cpp:935-1031"""
        F, g, pc, nrm = self.F, self.geom, self.pc, self.normalized
        T = self.T
        G = g.g
        v = (slice(G, -G), slice(G, -G))
        for b in self.beams:
            self.beam_slice(b, islice)
        for nm in ('ExmBy', 'EypBx', 'jx', 'jy', 'jz', 'rhomjz'):                  # Fields.cpp:565-566
            T(nm)[...] = 0.0
        if self.deposit_rho:
            T('rho')[...] = 0.0
        for pl in self.plasmas:                                                   # :617-618
            self.n_qsa_violation += deposit_current(
                pl, F, g, pc, nrm, jx=T('jx'), jy=T('jy'), jz=T('jz'),
                rho=T('rho') if self.deposit_rho else None, rhomjz=T('rhomjz'))
        for b in self.beams:                                                      # :621-623
            beam_deposit(self.beam_slice(b, islice), b, g, pc, nrm,
                         jxb=T('jx') if self.do_beam_jx_jy else None,
                         jyb=T('jy') if self.do_beam_jx_jy else None, jzb=T('jz'))
        if self.any_neutral:                                                      # :626
            T('rhomjz')[...] += F[('RhomJzIons', 'rhomjz')]
            if self.deposit_rho:
                T('rho')[...] += F[('RhomJzIons', 'rhomjz')]
        if self.grid_current is not None:                                         # :629 (jz, GridCurrent.cpp:48-49)
            grid_current_deposit(T('jz'), g, islice, *self.grid_current)
        solve_poisson_psi_ez_bz(F, g, pc, self.eig, self.open_bc)                 # :633

        def rel_error(a, b):                                                      # Fields.cpp:1227-1286
            ax, ay, bx, by = F[(a, 'Bx')][v], F[(a, 'By')][v], F[(b, 'Bx')][v], F[(b, 'By')][v]
            norm_b = np.sqrt(ax * ax + ay * ay).sum()
            norm_d = np.sqrt((ax - bx) * (ax - bx) + (ay - by) * (ay - by)).sum()
            return norm_d / norm_b if norm_b > 0.0 else 0.0

        err = rel_error('Previous', 'PCPrevIter')                                 # :941-943
        mix0 = math.exp(-0.5 * (err / (2.5 * self.predcorr_tol)) ** 2)            # Fields.cpp:1149-1170
        for c in ('Bx', 'By'):
            T(c)[...] = (1.0 + mix0) * F[('Previous', c)] + (-mix0) * F[('PCPrevIter', c)]
            F[('PCIter', c)][...] = 0.0                                           # :951-955
            F[('PCPrevIter', c)][...] = T(c)
        bc = (lambda r: open_boundary_rhs(r, g, True)) if self.open_bc else (lambda r: r)
        i_iter, err, err_prev = 0, 1.0, 1.0
        while err > self.predcorr_tol and i_iter < self.predcorr_max_iter:        # :961-1010
            i_iter += 1
            for pl in self.plasmas:                                               # push to the temp slice
                advance_plasma_particles(pl, F, g, pc, self.bc_kind, self.bc_lo, self.bc_hi, temp_slice=True)
            for pl in self.plasmas:                                               # jx jy of the next slice
                self.n_qsa_violation += deposit_current(pl, F, g, pc, nrm, jx=F[('Next', 'jx')],
                                                        jy=F[('Next', 'jy')])
            if self.do_beam_jx_jy and islice - 1 >= 0:
                for b in self.beams:
                    beam_deposit(self.beam_slice(b, islice - 1), b, g, pc, nrm,
                                 jxb=F[('Next', 'jx')], jyb=F[('Next', 'jy')])
            # SolvePoissonBxBy, Fields.cpp:1008-1078 -> PCIter
            dz_jy = (F[('Previous', 'jy')][v] - F[('Next', 'jy')][v]) * (0.5 / g.dz)
            dz_jx = (F[('Previous', 'jx')][v] - F[('Next', 'jx')][v]) * (0.5 / g.dz)
            F[('PCIter', 'Bx')][v] = poisson_dirichlet(
                bc(-pc.mu0 * _ddy(T('jz'), g.dy, G) + pc.mu0 * dz_jy), self.eig)
            F[('PCIter', 'By')][v] = poisson_dirichlet(
                bc(pc.mu0 * _ddx(T('jz'), g.dx, G) + (-pc.mu0) * dz_jx), self.eig)
            err = rel_error('This', 'PCIter')
            if i_iter == 1:
                err_prev = err
            # MixAndShiftBfields, Fields.cpp:1172-1225
            if err != 0.0 or err_prev != 0.0:
                w_it, w_prev = err_prev / (err + err_prev), err / (err + err_prev)
            else:
                w_it = w_prev = 0.5
            for c in ('Bx', 'By'):
                F[('PCPrevIter', c)][...] = w_it * F[('PCIter', c)] + w_prev * F[('PCPrevIter', c)]
                T(c)[...] = (1.0 - self.predcorr_mix) * T(c) + self.predcorr_mix * F[('PCPrevIter', c)]
                F[('PCPrevIter', c)][...] = F[('PCIter', c)]
            F[('Next', 'jx')][...] = 0.0                                          # :996-999
            F[('Next', 'jy')][...] = 0.0
            err_prev = err
        self.predcorr_iters.append(i_iter)
        if self.slice_hook:
            self.slice_hook(self, islice, 'fields')
        for b in self.beams:                                                      # :682-683 (before the push)
            bs = self.beam_slice(b, islice)
            n = bs['np']
            self.beam_diag.setdefault(b.name, []).append(
                {k: bs[k][:n].copy() for k in ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz', 'id')})
        self._accumulate_checksums(islice)
        for pl in self.plasmas:
            self.n_plasma_pushed += pl.x.size
            advance_plasma_particles(pl, F, g, pc, self.bc_kind, self.bc_lo, self.bc_hi)
        for b in self.beams:
            bs = self.beam_slice(b, islice)
            self.n_beam_pushed += bs['np']
            advance_beam_slice(bs, b, F, g, pc, islice, self.dt, getattr(self, 'time', 0.0),
                               self.bc_kind, self.bc_lo, self.bc_hi)
            shift_slipped_particles(bs, self.beam_slice(b, islice - 1), g, islice)
        if self.adaptive_dt:                                                      # :715
            self._adaptive_gather(islice)
        if self.slice_hook:
            self.slice_hook(self, islice, 'pushed')
        for c in ('Bx', 'By'):                                                    # ShiftSlices, Fields.cpp:600-603
            F[('PCPrevIter', c)][...] = F[('Previous', c)]
            F[('Previous', c)][...] = T(c)
        F[('Previous', 'jx')][...] = T('jx')
        F[('Previous', 'jy')][...] = T('jy')

    def solve_one_slice(self, islice: int):
        """Hipace::SolveOneSlice, Hipace.cpp:556-728 (explicit branch)."""
        for pl in self.plasmas:                                                   # :587
            per = self.plasma_insitu_period.get(pl.name, 0)
            if per > 0 and (self.step == self.max_step or self.step % per == 0):
                self.plasma_insitu.setdefault(pl.name, np.zeros((15, self.geom.nz)))[:, islice] = \
                    plasma_insitu_sums(pl, self.pc, self.plasma_insitu_radius.get(pl.name, math.inf))
        if not self.explicit:
            return self._solve_one_slice_pc(islice)
        F, g, pc, nrm = self.F, self.geom, self.pc, self.normalized
        T = self.T
        for b in self.beams:
            self.beam_slice(b, islice)
        # InitializeSlices, fields/Fields.cpp:535-586
        for nm in ('chi', 'Sy', 'Sx', 'ExmBy', 'EypBx', 'jz_beam', 'rhomjz'):
            T(nm)[...] = 0.0
        F[('Next', 'jx_beam')][...] = 0.0
        F[('Next', 'jy_beam')][...] = 0.0
        if self.deposit_rho:
            T('rho')[...] = 0.0
        aabs = None
        if self.use_laser:                                        # get_data :583 / :640, UpdateLaserAabs :603
            if self.step == 0:
                self.L.n00j00 = laser_envelope_slice(self.lasers, self.laser_lambda0, g, islice)
            else:                                                 # MultiBuffer::unpack_data :913-923
                self.L.n00j00, self.L.nm1j00 = (a.copy() for a in self.laser_store[islice])
            env = self.L.n00j00
            update_laser_aabs(env, T('aabs'), g, self.laser_interp_order)
            aabs = T('aabs')
            self.checksums['laserEnvelope'] = self.checksums.get('laserEnvelope', 0.0) \
                + float(np.abs(self._diag_rows(env, 0)).sum())
        # plasma deposit (:609-610)
        for pl in self.plasmas:
            self.n_qsa_violation += deposit_current(
                pl, F, g, pc, nrm, jx=T('jx'), jy=T('jy'),
                rho=T('rho') if self.deposit_rho else None, chi=T('chi'), rhomjz=T('rhomjz'),
                aabs=aabs)
        # beam deposit on This: jz_beam (:613-614)
        for b in self.beams:
            beam_deposit(self.beam_slice(b, islice), b, g, pc, nrm, jzb=T('jz_beam'))
        # AddRhoIons, fields/Fields.cpp:606-615
        if self.any_neutral:
            T('rhomjz')[...] += F[('RhomJzIons', 'rhomjz')]
            if self.deposit_rho:
                T('rho')[...] += F[('RhomJzIons', 'rhomjz')]
        if self.grid_current is not None:                                         # :629
            grid_current_deposit(T('jz_beam'), g, islice, *self.grid_current)
        if self.slice_hook:
            self.slice_hook(self, islice, 'deposited')
        solve_poisson_psi_ez_bz(F, g, pc, self.eig, False, self.field_periodic, self.poisson_periodic)   # :633
        if self.use_laser and self.dt != 0.0:                                     # :637 AdvanceSlice
            chi = laser_interpolate_chi(T('chi'), self.laser_chi_initial, g, self.laser_interp_order)
            if self.laser_solver == 'fft':
                laser_advance_fft(self.L, chi, g, pc, self.laser_lambda0, self.dt, self.step,
                                  self.laser_use_phase)
            else:
                if self.laser_mg is None:
                    self.laser_mg = MultiGrid2(g.dx, g.dy, g.nx, g.ny)
                laser_advance_mg(self.L, chi, g, pc, self.laser_lambda0, self.dt, self.step,
                                 self.laser_mg, self.laser_use_phase, self.laser_mg_avg_rhs,
                                 self.laser_mg_tol_rel, self.laser_mg_tol_abs)
            # MultiBuffer::pack_data :840-851: A^{n+1} and A^n of this slice go to the next step
            self.laser_next[islice] = (self.L.np1j00.copy(), self.L.n00j00.copy())
        # Next-slice beam jx/jy (:639-657)
        if self.do_beam_jx_jy:
            for b in self.beams:
                if islice - 1 >= 0:
                    beam_deposit(self.beam_slice(b, islice - 1), b, g, pc, nrm,
                                 jxb=F[('Next', 'jx_beam')], jyb=F[('Next', 'jy_beam')])
        init_sxsy_with_beam(F, g, pc)                                             # :660
        for pl in self.plasmas:
            explicit_deposition(pl, F, g, pc, nrm, aabs=aabs)                     # :663
        if self.slice_hook:
            self.slice_hook(self, islice, 'sources')
        # ExplicitMGSolveBxBy, Hipace.cpp:793-933
        v = (slice(g.g, -g.g), slice(g.g, -g.g))
        sol = np.stack([T('Bx')[v], T('By')[v]])
        if self.field_periodic:                                                   # Hipace.cpp:817-821
            enforce_periodic([T('Sy'), T('Sx'), T('chi')], g.g, True)
        rhs = np.stack([T('Sy')[v], T('Sx')[v]])
        self.mg.solve1(sol, rhs, T('chi')[v], self.mg_tol_rel, self.mg_tol_abs, 200)
        T('Bx')[v], T('By')[v] = sol[0], sol[1]
        if self.field_periodic:                                                   # Hipace.cpp:924-927
            enforce_periodic([T('Bx'), T('By')], g.g, False)
        self.mg_cycles.append(self.mg.n_vcycles_last)
        if self.slice_hook:
            self.slice_hook(self, islice, 'fields')
        if self.field_insitu_period > 0 and (self.step == self.max_step
                                             or self.step % self.field_insitu_period == 0):   # :685
            if self.field_insitu is None:
                self.field_insitu = np.zeros((10, g.nz))
            self.field_insitu[:, islice] = field_insitu_sums(F, g, pc)
        if self.laser_insitu_period > 0 and (self.step == self.max_step
                                             or self.step % self.laser_insitu_period == 0):   # :688
            if self.laser_insitu is None:
                self.laser_insitu = np.zeros((8, g.nz))
            self.laser_insitu[:, islice] = laser_insitu_sums(self.L.n00j00, g)
        self._accumulate_checksums(islice)                                        # :691
        for b in self.beams:                                                      # :682-683 (before the push)
            bs = self.beam_slice(b, islice)
            n = bs['np']
            self.beam_diag.setdefault(b.name, []).append(
                {k: bs[k][:n].copy() for k in ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz', 'id')})
        for pl in self.plasmas:                                                   # :699-701
            self.n_plasma_pushed += pl.x.size
            advance_plasma_particles(pl, F, g, pc, self.bc_kind, self.bc_lo, self.bc_hi, aabs=aabs)
        for b in self.beams:                                                      # :680-681
            per = self.insitu_period.get(b.name, 0)
            if per > 0 and (self.step == self.max_step or self.step % per == 0):
                self.insitu.setdefault(b.name, np.zeros((23, g.nz)))[:, islice] = \
                    beam_insitu_sums(self.beam_slice(b, islice), pc, self.insitu_radius.get(b.name, math.inf))
        for b in self.beams:                                                      # :707-709
            bs = self.beam_slice(b, islice)
            self.n_beam_pushed += bs['np']
            advance_beam_slice(bs, b, F, g, pc, islice, self.dt, getattr(self, 'time', 0.0),
                               self.bc_kind, self.bc_lo, self.bc_hi)
            shift_slipped_particles(bs, self.beam_slice(b, islice - 1), g, islice)
        if self.adaptive_dt:                                                      # :715
            self._adaptive_gather(islice)
        if self.slice_hook:
            self.slice_hook(self, islice, 'pushed')
        # ShiftSlices, fields/Fields.cpp:588-604
        F[('Previous', 'jx_beam')][...] = T('jx_beam')
        F[('Previous', 'jy_beam')][...] = T('jy_beam')
        T('jx_beam')[...] = F[('Next', 'jx_beam')]
        T('jy_beam')[...] = F[('Next', 'jy_beam')]
        T('jx')[...] = F[('Next', 'jx_beam')]
        T('jy')[...] = F[('Next', 'jy_beam')]
        if self.use_laser:
            self.L.shift()                                                        # :727

    def _diag_rows(self, a, g):
        """field diagnostic of one slice (fields/Fields.cpp:413-533): the valid box for 'xyz'; for
        'xz' the order-1 interpolation to y = mid-domain (diagnostics/Diagnostic.cpp:393-407),
        i.e. the mean of the two central rows (even ny) or the central row.  g = guard width of a."""
        ny = a.shape[0] - 2 * g
        v = a[g:g + ny, g:a.shape[1] - g]
        if self.diag_type == 'xyz':
            return v
        return 0.5 * (v[ny // 2 - 1] + v[ny // 2]) if ny % 2 == 0 else v[ny // 2]

    def _accumulate_checksums(self, islice):
        """checksum of a field = sum|Q| over the last iteration's valid cells,
        tests/checksum/backend/openpmd_backend.py:40-45 (diag copy fields/Fields.cpp:413-533
        is the identity for an uncoarsened xyz diagnostic)."""
        for (sl, nm), a in self.F.items():
            if sl == 'This':
                self.checksums[nm] = self.checksums.get(nm, 0.0) + float(np.abs(self._diag_rows(a, self.geom.g)).sum())

    def end_step_adaptive(self, step):
        """Hipace.cpp:482-483: after the last slice the owning rank derives its next dt from the beam"""
        if self.adaptive_dt:
            self._adaptive_from_min_uz(self.time)
            self._rank_state[step % self.numprocs] = (self.dt, self._min_uz_mq)

    def evolve(self, nslices: int | None = None, step_begin: int = 0, step_end: int = 0):
        """Run time steps step_begin..step_end (Hipace.cpp:401-507); the checksums returned are
        those of the last one.  nslices limits the slice loop (from the head) for bounded
        tests/benchmarks."""
        g = self.geom
        stop = -1 if nslices is None else max(-1, g.nz - 1 - nslices)
        for step in range(step_begin, step_end + 1):
            self.checksums = {}
            self.begin_step(step)
            self.insitu, self.plasma_insitu, self.field_insitu, self.laser_insitu = {}, {}, None, None
            for isl in range(g.nz - 1, stop, -1):
                self.solve_one_slice(isl)
            if stop == -1:
                self.end_step_adaptive(step)
            for b in self.beams:                                                  # Hipace.cpp:488
                if b.name in self.insitu:
                    ndf = g.dx * g.dy * g.dz if self.normalized else 1.0
                    self.insitu_records.setdefault(b.name, []).append(insitu_beam_record(
                        self.insitu[b.name], self.time, step, b.charge, b.mass, g.lo[2], g.hi[2],
                        ndf, self.normalized)[1])
            if self.field_insitu is not None:                                     # Hipace.cpp:487
                self.field_insitu_records.append(insitu_field_record(
                    self.field_insitu, self.time, step, g.lo[2], g.hi[2], self.normalized,
                    g.dx * g.dy * g.dz)[1])
            if self.laser_insitu is not None:                                     # Hipace.cpp:490
                self.laser_insitu_records.append(insitu_laser_record(
                    self.laser_insitu, self.time, step, g.lo[2], g.hi[2], self.normalized,
                    g.dx * g.dy * g.dz, g.nx, g.ny)[1])
            for pl in self.plasmas:                                               # Hipace.cpp:489
                if pl.name in self.plasma_insitu:
                    ndf = g.dx * g.dy * g.dz if self.normalized else 1.0
                    self.plasma_insitu_records.setdefault(pl.name, []).append(insitu_plasma_record(
                        self.plasma_insitu[pl.name], self.time, step, pl.charge, pl.mass, g.lo[2], g.hi[2],
                        ndf, self.normalized)[1])
            if self.use_laser:
                self.laser_store = self.laser_next
        return self.checksums

    def beam_checksums(self):
        out = {}
        for b in self.beams:
            diag = self.beam_diag.get(b.name, [])
            if len(diag) == self.geom.nz:      # state written by the last step's diagnostics
                cat = lambda k, _d=diag: np.concatenate([t[k] for t in _d])
            else:
                for isl in range(self.geom.nz):
                    self.beam_slice(b, isl)
                cat = lambda k: np.concatenate([b.slices[i][k][:b.slices[i]['np']] for i in sorted(b.slices)])
            n = cat('x').size
            # momenta are stored as proper velocities u c (BeamParticleContainerInit.cpp:52-61); the
            # checksum reads them through openPMD-viewer, which normalises momentum to m c
            ci = 1.0 / self.pc.c
            out[b.name] = dict(
                charge=abs(b.charge) * n, mass=abs(b.mass) * n, id=int(np.abs(cat('id')).sum()),
                x=float(np.abs(cat('x')).sum()), y=float(np.abs(cat('y')).sum()),
                z=float(np.abs(cat('z')).sum()), ux=float(np.abs(cat('ux')).sum()) * ci,
                uy=float(np.abs(cat('uy')).sum()) * ci, uz=float(np.abs(cat('uz')).sum()) * ci,
                w=float(np.abs(cat('w')).sum()))
        return out
