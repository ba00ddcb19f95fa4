"""Multi-threaded CPU restatement: oracle.hipace_oracle.Simulation with its kernels replaced by
the C/OpenMP functions of oracle/hipace_cport.c (deposit, explicit deposit, gather+push, field
stencils, hpmg) and SciPy's pocketfft DST (workers = all cores) for the Poisson solves -- the
stand-in for FFTW RODFT00, the reference's CPU default (src/fields/Fields.cpp:37-38).

THIS IS TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE (see the header of hipace_oracle.py).
Used by bench.py's cpu_baseline and --impl reference legs.  Pinned by tests/test_cport.py
against the NumPy oracle per cell and against the reference's golden checksums.

Build: gcc -O3 -march=native -fopenmp -shared -fPIC (make_cport()), output oracle/_ref/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import hipace_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'hipace_cport.c')
_OUT = os.path.join(_HERE, '_ref')
_LIB = os.path.join(_OUT, 'libhipace_cport.so')
G = O.G

try:
    from scipy.fft import dstn as _dstn
except Exception:  # pragma: no cover
    _dstn = None


def _host_stamp():
    """source hash + CPU model: -march=native code built on one host must not run on another"""
    import hashlib
    h = hashlib.sha256(open(_SRC, 'rb').read())
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith(('model name', 'flags')):
                h.update(line.encode())
                if line.startswith('flags'):
                    break
    except OSError:
        pass
    return h.hexdigest()


def make_cport(force=False):
    os.makedirs(_OUT, exist_ok=True)
    stamp_file = _LIB + '.stamp'
    stamp = _host_stamp()
    if (not force and os.path.exists(_LIB) and os.path.exists(stamp_file)
            and open(stamp_file).read() == stamp):
        return _LIB
    cmd = ['gcc', '-O3', '-march=native', '-fopenmp', '-ffp-contract=off', '-shared', '-fPIC',
           '-o', _LIB, _SRC, '-lm']
    subprocess.run(cmd, check=True)
    open(stamp_file, 'w').write(stamp)
    return _LIB


_lib = None
_dp = C.POINTER(C.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        make_cport()
        L = C.CDLL(_LIB)
        L.hpc_deposit_current.restype = C.c_long
        L.hpc_deposit_current.argtypes = [C.c_long] + [C.c_void_p] * 12 + [C.c_int, C.c_int] + [C.c_double] * 8
        L.hpc_explicit_deposition.restype = None
        L.hpc_explicit_deposition.argtypes = [C.c_long] + [C.c_void_p] * 13 + [C.c_int, C.c_int] + [C.c_double] * 8
        L.hpc_advance_plasma.restype = None
        L.hpc_advance_plasma.argtypes = ([C.c_long] + [C.c_void_p] * 17 + [C.c_int, C.c_int]
                                         + [C.c_double] * 7 + [C.c_int] * 3 + [C.c_double] * 4)
        L.hpc_poisson_rhs.restype = None
        L.hpc_poisson_rhs.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_void_p] + [C.c_double] * 5
        L.hpc_store_valid.restype = None
        L.hpc_store_valid.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.hpc_exmby_eypbx.restype = None
        L.hpc_exmby_eypbx.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_double, C.c_double]
        L.hpc_sxsy_from_beam.restype = None
        L.hpc_sxsy_from_beam.argtypes = [C.c_void_p] * 7 + [C.c_int, C.c_int] + [C.c_double] * 4
        L.hpc_mg_create.restype = C.c_void_p
        L.hpc_mg_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.hpc_mg_destroy.argtypes = [C.c_void_p]
        L.hpc_mg_solve1.restype = C.c_int
        L.hpc_mg_solve1.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
        L.hpc_num_threads.restype = C.c_int
        L.hpc_set_num_threads.restype = C.c_int
        L.hpc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def physical_cores():
    """physical cores this process may run on (the reference's `nosmt` policy, Parser.H:80-82)"""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    try:
        import psutil
        phys = psutil.cpu_count(logical=False) or avail
    except Exception:
        phys = avail
    return max(1, min(avail, phys))


_threads_override = None


def set_threads(n):
    """force the thread count of the C kernels AND of SciPy's DST (workers), whatever
    OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)"""
    global _threads_override
    _threads_override = lib().hpc_set_num_threads(int(n))
    return _threads_override


class Simulation(O.Simulation):
    """Same driver (Hipace.cpp:393-728 restated in hipace_oracle.Simulation), C/OpenMP kernels."""

    def __init__(self, deck_text, overrides=None):
        super().__init__(deck_text, overrides)
        self.L = lib()
        self.threads = self.L.hpc_num_threads()
        g = self.geom
        # one contiguous slice array so that adjacent components (Bx,By), (Sy,Sx) are adjacent
        # in memory as in the reference (Hipace.cpp:806-810)
        names = list(self.F.keys())
        self._arr = np.zeros((len(names), g.ny + 2 * G, g.nx + 2 * G))
        self.F = {k: self._arr[i] for i, k in enumerate(names)}
        self._mg = self.L.hpc_mg_create(g.nx, g.ny, g.dx, g.dy)
        assert self._mg, 'hpmg: nx and ny must have the same parity'
        self._rhs = np.zeros((3, g.ny, g.nx))
        self._xoff, self._yoff = g.pos_offset(0), g.pos_offset(1)

    def __del__(self):
        try:
            self.L.hpc_mg_destroy(self._mg)
        except Exception:
            pass

    def first_beam_slot(self):
        """number of (empty) slices ahead of the first beam slice, counted from the head"""
        g = self.geom
        zmax = max((b.zmax for b in self.beams), default=g.hi[2])
        n = int(np.floor((g.hi[2] - zmax) / g.dz - 0.5)) + 1
        return min(max(n, 0), g.nz - 1)

    # -- kernels ---------------------------------------------------------------------------
    def _deposit(self, pl, jx, jy, rho, chi, rhomjz, flip=False):
        g, pc = self.geom, self.pc
        charge = -pl.charge if flip else pl.charge
        invvol = 1.0 if self.normalized else 1.0 / (g.dx * g.dy * g.dz)
        v8 = pl.valid.view(np.uint8)
        return self.L.hpc_deposit_current(
            pl.x.size, _p(pl.x), _p(pl.y), _p(pl.w), _p(pl.ux), _p(pl.uy), _p(pl.psi), _p(v8),
            _p(jx), _p(jy), _p(rho), _p(chi), _p(rhomjz), g.nx, g.ny, self._xoff, self._yoff,
            1.0 / g.dx, 1.0 / g.dy, 1.0 / pc.c, charge * invvol, charge * pc.mu0 / pl.mass,
            pl.max_qsa_weighting_factor)

    def begin_step(self, step=0):
        self.step = step
        self.time = step * self.dt
        self.beam_diag = {b.name: [] for b in self.beams}
        self._arr[...] = 0.0
        for b in self.beams:
            for bs in b.slices.values():
                bs['nsub'][...] = 0
        for pl in self.plasmas:
            O.init_plasma(pl, self.geom, self.pc, self.normalized, self.bc_lo, self.bc_hi,
                          c_t=self.pc.c * self.time)
        for pl in self.plasmas:
            if pl.neutralize_background:
                self._deposit(pl, None, None, None, None, self.F[('RhomJzIons', 'rhomjz')], flip=True)

    def solve_one_slice(self, islice):
        """Hipace::SolveOneSlice (Hipace.cpp:556-728), explicit branch; same order as the NumPy
        oracle's solve_one_slice."""
        F, g, pc, nrm, L = self.F, self.geom, self.pc, self.normalized, self.L
        T = self.T
        for nm in ('chi', 'Sy', 'Sx', 'ExmBy', 'EypBx', 'jz_beam', 'rhomjz'):
            T(nm)[...] = 0.0
        F[('Next', 'jx_beam')][...] = 0.0
        F[('Next', 'jy_beam')][...] = 0.0
        if self.deposit_rho:
            T('rho')[...] = 0.0
        for pl in self.plasmas:
            self.n_qsa_violation += self._deposit(pl, T('jx'), T('jy'),
                                                  T('rho') if self.deposit_rho else None,
                                                  T('chi'), T('rhomjz'))
        for b in self.beams:
            O.beam_deposit(self.beam_slice(b, islice), b, g, pc, nrm, jzb=T('jz_beam'))
        if self.any_neutral:
            T('rhomjz')[...] += F[('RhomJzIons', 'rhomjz')]
            if self.deposit_rho:
                T('rho')[...] += F[('RhomJzIons', 'rhomjz')]
        # three Poisson solves (Fields.cpp:880-918)
        L.hpc_poisson_rhs(_p(T('rhomjz')), _p(T('jx')), _p(T('jy')), g.nx, g.ny, _p(self._rhs),
                          -1.0 / pc.ep0, 1.0 / (pc.ep0 * pc.c), pc.mu0, g.dx, g.dy)
        spec = _dstn(self._rhs, type=1, axes=(1, 2), workers=(_threads_override or -1))
        spec *= self.eig
        sol = _dstn(spec, type=1, axes=(1, 2), workers=(_threads_override or -1))
        for k, nm in enumerate(('Psi', 'Ez', 'Bz')):
            L.hpc_store_valid(_p(T(nm)), _p(sol[k]), g.nx, g.ny)
        L.hpc_exmby_eypbx(_p(T('Psi')), _p(T('ExmBy')), _p(T('EypBx')), g.nx, g.ny, g.dx, g.dy)
        if self.do_beam_jx_jy:
            for b in self.beams:
                if islice - 1 >= 0:
                    O.beam_deposit(self.beam_slice(b, islice - 1), b, g, pc, nrm,
                                   jxb=F[('Next', 'jx_beam')], jyb=F[('Next', 'jy_beam')])
        L.hpc_sxsy_from_beam(_p(T('Sy')), _p(T('Sx')), _p(T('jz_beam')),
                             _p(F[('Previous', 'jx_beam')]), _p(F[('Previous', 'jy_beam')]),
                             _p(F[('Next', 'jx_beam')]), _p(F[('Next', 'jy_beam')]),
                             g.nx, g.ny, pc.mu0, g.dx, g.dy, g.dz)
        invvol = 1.0 if nrm else 1.0 / (g.dx * g.dy * g.dz)
        for pl in self.plasmas:
            L.hpc_explicit_deposition(
                pl.x.size, _p(pl.x), _p(pl.y), _p(pl.w), _p(pl.ux), _p(pl.uy), _p(pl.psi),
                _p(pl.valid.view(np.uint8)), _p(T('Sy')), _p(T('Sx')), _p(T('Bz')), _p(T('Ez')),
                _p(T('ExmBy')), _p(T('EypBx')), g.nx, g.ny, self._xoff, self._yoff, 1.0 / g.dx,
                1.0 / g.dy, pc.c, 1.0 / pc.c, pl.charge * invvol * pc.mu0, pl.charge / pl.mass)
        if self.slice_hook:
            self.slice_hook(self, islice, 'sources')
        # (Bx,By) and (Sy,Sx) are adjacent components of self._arr
        it = L.hpc_mg_solve1(self._mg, _p(T('Bx')), _p(T('Sy')), _p(T('chi')), g.nx, g.ny,
                             self.mg_tol_rel, self.mg_tol_abs, 200)
        if it < 0:
            raise RuntimeError('hpmg failed')
        self.mg_cycles.append(it)
        if self.slice_hook:
            self.slice_hook(self, islice, 'fields')
        self._accumulate_checksums(islice)
        for b in self.beams:
            bs = self.beam_slice(b, islice)
            n = bs['np']
            self.beam_diag.setdefault(b.name, []).append(
                {k: bs[k][:n].copy() for k in ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz', 'id')})
        bc = {'Reflecting': 0, 'Periodic': 1, 'Absorbing': 2}[self.bc_kind]
        for pl in self.plasmas:
            self.n_plasma_pushed += pl.x.size
            L.hpc_advance_plasma(
                pl.x.size, _p(pl.x), _p(pl.y), _p(pl.w), _p(pl.ux), _p(pl.uy), _p(pl.psi),
                _p(pl.x_prev), _p(pl.y_prev), _p(pl.ux_half), _p(pl.uy_half), _p(pl.psi_half),
                _p(pl.valid.view(np.uint8)), _p(T('Psi')), _p(T('Ez')), _p(T('Bx')), _p(T('By')),
                _p(T('Bz')), g.nx, g.ny, self._xoff, self._yoff, 1.0 / g.dx, 1.0 / g.dy, pc.c,
                pl.charge / (pl.mass * pc.c), g.dz / pl.n_subcycles, pl.n_subcycles, 0, bc,
                self.bc_lo[0], self.bc_lo[1], self.bc_hi[0], self.bc_hi[1])
        # the beam slice is ~1e-3 of the plasma work: NumPy restatement (hipace_oracle.py)
        for b in self.beams:
            bs = self.beam_slice(b, islice)
            self.n_beam_pushed += bs['np']
            O.advance_beam_slice(bs, b, F, g, pc, islice, self.dt, self.time, self.bc_kind,
                                 self.bc_lo, self.bc_hi)
            O.shift_slipped_particles(bs, self.beam_slice(b, islice - 1), g, islice)
        if self.slice_hook:
            self.slice_hook(self, islice, 'pushed')
        F[('Previous', 'jx_beam')][...] = T('jx_beam')
        F[('Previous', 'jy_beam')][...] = T('jy_beam')
        T('jx_beam')[...] = F[('Next', 'jx_beam')]
        T('jy_beam')[...] = F[('Next', 'jy_beam')]
        T('jx')[...] = F[('Next', 'jx_beam')]
        T('jy')[...] = F[('Next', 'jy_beam')]

    @property
    def nz(self):
        return self.geom.nz
