"""hipace_b200 -- host-side Python mirror of the HiPACE++ slice-loop call surface over libhpb200.so.

The product is the C-ABI library (include/hpb200.h, hipace_b200/csrc/*.cu); this module is a thin
ctypes binding with the names of the reference seams it replaces:

  Simulation(deck, overrides)        Hipace::Hipace + InitData          src/Hipace.cpp:74-295
  Simulation.evolve()                Hipace::Evolve                     src/Hipace.cpp:393-554
  Simulation.solve_one_slice(isl)    Hipace::SolveOneSlice              src/Hipace.cpp:556-728
  Context.deposit_current(...)       ::DepositCurrent                   PlasmaDepositCurrent.cpp:22
  Context.explicit_deposition(...)   ::ExplicitDeposition               ExplicitDeposition.cpp:20
  Context.advance_plasma_particles   AdvancePlasmaParticles             PlasmaParticleAdvance.cpp:29
  Context.poisson_solve(...)         FFTPoissonSolver::SolvePoissonEquation
  Context.mg_solve1(...)             hpmg::MultiGrid::solve1            HpMultiGrid.cpp:1169
  Context.advance_beam_particles     AdvanceBeamParticlesSlice          BeamParticleAdvance.cpp:19
  Context.beam_shift_slipped         shiftSlippedParticles              SliceSort.cpp:13
  Simulation.pipeline_init(...)      MultiBuffer::initialize            MultiBuffer.cpp (NCCL p2p)

There is NO CPU fallback: importing works without a GPU (so the symbol table can be checked), but
every compute entry point needs a CUDA device and raises HpbError otherwise, and a missing
libhpb200.so raises at import of the library handle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libhpb200.so')

NGUARD = 2
PLASMA_REALS = ('x', 'y', 'w', 'ux', 'uy', 'psi', 'x_prev', 'y_prev', 'ux_half_step',
                'uy_half_step', 'psi_half_step')
BC = {'Reflecting': 0, 'Periodic': 1, 'Absorbing': 2}

# enum hpb_comp
COMP_IDS = ('NEXT_JX_BEAM', 'NEXT_JY_BEAM', 'CHI', 'SY', 'SX', 'EXMBY', 'EYPBX', 'EZ', 'BX', 'BY',
            'BZ', 'PSI', 'JX_BEAM', 'JY_BEAM', 'JZ_BEAM', 'JX', 'JY', 'RHOMJZ', 'RHO',
            'PREV_JX_BEAM', 'PREV_JY_BEAM', 'IONS_RHOMJZ', 'AABS',
            'JZ', 'NEXT_JX', 'NEXT_JY', 'PREV_BX', 'PREV_BY', 'PREV_JX', 'PREV_JY',
            'PCITER_BX', 'PCITER_BY', 'PCPREV_BX', 'PCPREV_BY')
COMP = {n: i for i, n in enumerate(COMP_IDS)}

# every symbol include/hpb200.h declares (tests check the library exports all of them)
EXPORTS = (
    'hpb_create', 'hpb_destroy', 'hpb_last_error', 'hpb_version', 'hpb_deposit_current',
    'hpb_beam_deposit', 'hpb_fields_initialize_slices', 'hpb_fields_add_rho_ions',
    'hpb_fields_shift_slices', 'hpb_poisson_solve', 'hpb_fields_solve_psi_ez_bz',
    'hpb_fields_sxsy_from_beam', 'hpb_explicit_deposition', 'hpb_mg_solve1',
    'hpb_advance_plasma_particles', 'hpb_abs_sum', 'hpb_sim_create', 'hpb_sim_destroy',
    'hpb_sim_evolve', 'hpb_sim_begin_step', 'hpb_sim_solve_one_slice', 'hpb_sim_geometry',
    'hpb_sim_ncomp', 'hpb_sim_comp_index', 'hpb_sim_get_field', 'hpb_sim_set_field',
    'hpb_sim_plasma_np', 'hpb_sim_get_plasma_real', 'hpb_sim_get_plasma_valid',
    'hpb_sim_checksum_count', 'hpb_sim_checksum_name', 'hpb_sim_get_checksums',
    'hpb_sim_get_beam_checksums', 'hpb_sim_get_stats', 'hpb_sim_set_option', 'hpb_sim_beam_np',
    'hpb_sim_get_beam', 'hpb_sim_set_beam', 'hpb_extfields_create', 'hpb_extfields_destroy',
    'hpb_advance_beam_particles', 'hpb_beam_shift_slipped', 'hpb_nccl_unique_id',
    'hpb_sim_pipeline_init', 'hpb_sim_pipeline_message_bytes', 'hpb_sim_beam_slice_capacity',
    'hpb_sim_timer_start', 'hpb_sim_timer_stop', 'hpb_sim_get_beam_packet',
    'hpb_fields_shift_and_initialize', 'hpb_advance_plasma_particles_and_deposit', 'hpb_deck_check',
    'hpb_set_plasma_lattice_hint', 'hpb_deposit_current_laser', 'hpb_laser_update_aabs',
    'hpb_set_deposition_order', 'hpb_fields_grid_current', 'hpb_sim_nguard',
    'hpb_beam_insitu_slice', 'hpb_insitu_write_beam', 'hpb_fields_zero', 'hpb_deposit_current_jz',
    'hpb_fields_bxby_rhs', 'hpb_fields_psi_ez_bz_rhs', 'hpb_fields_open_boundary',
    'hpb_fields_rel_b_error', 'hpb_fields_lincomb2', 'hpb_beam_min_uz_slice', 'hpb_adaptive_dt_next',
    'hpb_abs_sum_xz', 'hpb_plasma_insitu_slice', 'hpb_insitu_write_plasma',
    'hpb_fields_insitu_slice', 'hpb_insitu_write_fields', 'hpb_debug_push_thread_map',
    'hpb_laser_state_create', 'hpb_laser_state_destroy', 'hpb_laser_begin_step', 'hpb_laser_get_slice',
    'hpb_laser_advance_slice', 'hpb_laser_shift_slices', 'hpb_laser_end_step',
    'hpb_laser_insitu_slice', 'hpb_insitu_write_laser', 'hpb_sim_get_mg_iters', 'hpb_set_option', 'hpb_plasma_reorder', 'hpb_measure_fp64_peak', 'hpb_mg_solve2', 'hpb_laser_set_solver',
    'hpb_laser_mg_vcycles', 'hpb_poisson_solve_periodic', 'hpb_fields_enforce_periodic', 'hpb_sim_get_time', 'hpb_mg_prepare_acf', 'hpb_abs_sum_multi', 'hpb_debug_fft_prime_table',
)
NCCL_ID_BYTES = 128


class HpbError(RuntimeError):
    pass


class hpb_slice(C.Structure):
    _fields_ = [('p', C.c_void_p), ('lo_x', C.c_int), ('lo_y', C.c_int), ('nx_tot', C.c_int),
                ('ny_tot', C.c_int), ('jstride', C.c_long), ('nstride', C.c_long),
                ('ncomp', C.c_int)]


class hpb_plasma(C.Structure):
    _fields_ = [('r', C.c_void_p * 11), ('idcpu', C.c_void_p), ('np', C.c_long)]


class hpb_beam_slice(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz', 'idcpu')] + \
               [('np', C.c_long), ('d_np', C.c_void_p)]


class hpb_geom(C.Structure):
    _fields_ = [('nx', C.c_int), ('ny', C.c_int), ('dx', C.c_double), ('dy', C.c_double),
                ('dz', C.c_double), ('x_off', C.c_double), ('y_off', C.c_double),
                ('c', C.c_double), ('ep0', C.c_double), ('mu0', C.c_double), ('q_e', C.c_double),
                ('m_e', C.c_double), ('normalized', C.c_int)]


class hpb_sim_stats(C.Structure):
    _fields_ = [('n_plasma_pushed', C.c_double), ('n_beam_pushed', C.c_double),
                ('n_cells_updated', C.c_double), ('slice_loop_ms', C.c_double),
                ('n_slices', C.c_long), ('n_mg_vcycles', C.c_long), ('n_qsa_violation', C.c_long),
                ('n_kernel_launches', C.c_long), ('ms_deposit', C.c_double),
                ('ms_poisson', C.c_double), ('ms_explicit', C.c_double), ('ms_mg', C.c_double),
                ('ms_push', C.c_double), ('ms_other', C.c_double), ('n_reorders', C.c_long), ('n_fused_slices', C.c_long)]


_lib = None


def lib():
    """The loaded C-ABI library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HpbError(f'{LIB_PATH} is missing: run `python -m hipace_b200.build` '
                       '(there is no CPU fallback)')
    L = C.CDLL(LIB_PATH)
    L.hpb_last_error.restype = C.c_char_p
    L.hpb_version.restype = C.c_char_p
    L.hpb_sim_checksum_name.restype = C.c_char_p
    L.hpb_sim_plasma_np.restype = C.c_long
    L.hpb_destroy.restype = None
    L.hpb_sim_destroy.restype = None
    L.hpb_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(hpb_geom), C.c_void_p]
    L.hpb_destroy.argtypes = [C.c_void_p]
    L.hpb_deposit_current.argtypes = [C.c_void_p, hpb_plasma, hpb_slice, C.c_double, C.c_double,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                      C.c_void_p]
    L.hpb_beam_deposit.argtypes = [C.c_void_p, hpb_beam_slice, hpb_slice, C.c_double, C.c_int,
                                   C.c_int, C.c_int]
    IP = C.POINTER(C.c_int)
    for f in ('hpb_fields_initialize_slices', 'hpb_fields_add_rho_ions', 'hpb_fields_shift_slices',
              'hpb_fields_solve_psi_ez_bz', 'hpb_fields_sxsy_from_beam'):
        getattr(L, f).argtypes = [C.c_void_p, hpb_slice, IP]
    L.hpb_poisson_solve.argtypes = [C.c_void_p, C.c_void_p, hpb_slice, IP, C.c_int]
    L.hpb_poisson_solve_periodic.argtypes = [C.c_void_p, C.c_void_p, hpb_slice, IP, C.c_int]
    L.hpb_fields_enforce_periodic.argtypes = [C.c_void_p, hpb_slice, C.c_int, IP, C.c_int]
    L.hpb_explicit_deposition.argtypes = [C.c_void_p, hpb_plasma, hpb_slice, C.c_double,
                                          C.c_double, IP]
    L.hpb_mg_solve1.argtypes = [C.c_void_p, hpb_slice, C.c_int, C.c_int, C.c_int, C.c_double,
                                C.c_double, C.c_int, IP]
    DP = C.POINTER(C.c_double)
    L.hpb_advance_plasma_particles.argtypes = [C.c_void_p, hpb_plasma, hpb_slice, C.c_double,
                                               C.c_double, C.c_int, C.c_int, C.c_int, DP, DP, IP]
    L.hpb_abs_sum.argtypes = [C.c_void_p, hpb_slice, C.c_int, C.c_void_p]
    L.hpb_sim_create.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_char_p, C.c_int]
    L.hpb_sim_destroy.argtypes = [C.c_void_p]
    L.hpb_sim_evolve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.hpb_sim_begin_step.argtypes = [C.c_void_p, C.c_int]
    L.hpb_sim_solve_one_slice.argtypes = [C.c_void_p, C.c_int]
    L.hpb_sim_geometry.argtypes = [C.c_void_p, IP, DP, DP]
    L.hpb_sim_ncomp.argtypes = [C.c_void_p]
    L.hpb_sim_comp_index.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.hpb_sim_get_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.hpb_sim_set_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.hpb_sim_plasma_np.argtypes = [C.c_void_p, C.c_int]
    L.hpb_sim_get_plasma_real.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.hpb_sim_get_plasma_valid.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.hpb_sim_checksum_count.argtypes = [C.c_void_p]
    L.hpb_sim_checksum_name.argtypes = [C.c_void_p, C.c_int]
    L.hpb_sim_get_checksums.argtypes = [C.c_void_p, C.c_void_p]
    L.hpb_sim_get_beam_checksums.argtypes = [C.c_void_p, C.c_int, DP]
    L.hpb_sim_beam_np.restype = C.c_long
    L.hpb_sim_beam_np.argtypes = [C.c_void_p, C.c_int]
    L.hpb_sim_get_beam.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
    L.hpb_sim_set_beam.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]
    L.hpb_sim_get_stats.argtypes = [C.c_void_p, C.POINTER(hpb_sim_stats)]
    L.hpb_sim_get_time.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.hpb_sim_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.hpb_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.hpb_measure_fp64_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.hpb_mg_solve2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                C.c_double, C.c_int, C.POINTER(C.c_int)]
    L.hpb_sim_get_mg_iters.restype = C.c_long
    L.hpb_sim_get_mg_iters.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_long]
    L.hpb_extfields_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_char_p)]
    L.hpb_extfields_destroy.argtypes = [C.c_void_p]
    L.hpb_extfields_destroy.restype = None
    L.hpb_advance_beam_particles.argtypes = [
        C.c_void_p, hpb_beam_slice, C.c_void_p, hpb_slice, C.c_double, C.c_double, C.c_int,
        C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, DP, DP, IP, C.c_void_p, C.c_void_p,
        C.c_void_p]
    L.hpb_beam_shift_slipped.argtypes = [
        C.c_void_p, hpb_beam_slice, C.c_void_p, C.c_double, C.c_void_p, hpb_beam_slice, C.c_void_p,
        hpb_beam_slice, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hpb_sim_timer_start.argtypes = [C.c_void_p]
    L.hpb_sim_timer_stop.argtypes = [C.c_void_p, DP]
    L.hpb_sim_get_beam_packet.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.hpb_nccl_unique_id.argtypes = [C.c_char_p]
    L.hpb_sim_pipeline_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
    L.hpb_sim_pipeline_message_bytes.argtypes = [C.c_void_p]
    L.hpb_sim_pipeline_message_bytes.restype = C.c_long
    L.hpb_sim_beam_slice_capacity.argtypes = [C.c_void_p, C.c_int]
    L.hpb_sim_beam_slice_capacity.restype = C.c_long
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise HpbError(f'{what} failed (code {rc}): {lib().hpb_last_error().decode()}')


def _overrides_text(overrides):
    if not overrides:
        return b''
    lines = []
    for k, v in overrides.items():
        if isinstance(v, (list, tuple)):
            v = ' '.join(str(t) for t in v)
        lines.append(f'{k} = {v}')
    return ('\n'.join(lines) + '\n').encode()


def insitu_write_beam(path, time, step, charge, mass, z_lo, z_hi, density_factor, normalized, sums):
    """hpb_insitu_write_beam (host only): append one in-situ record built from the raw per-slice
    sums[23, n_slices] to `path` in the reference's NumPy-structured format"""
    a = np.ascontiguousarray(sums, dtype=np.float64)
    assert a.ndim == 2 and a.shape[0] == 23
    L = lib()
    L.hpb_insitu_write_beam.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int] + [C.c_double] * 5 + \
        [C.c_int, C.c_void_p]
    _check(L.hpb_insitu_write_beam(str(path).encode(), time, step, a.shape[1], charge, mass, z_lo, z_hi,
                                   density_factor, int(normalized), a.ctypes.data), 'hpb_insitu_write_beam')


def insitu_write_plasma(path, time, step, charge, mass, z_lo, z_hi, density_factor, normalized, sums):
    """hpb_insitu_write_plasma (host only): raw per-slice sums[15, n_slices] -> one appended record"""
    a = np.ascontiguousarray(sums, dtype=np.float64)
    assert a.ndim == 2 and a.shape[0] == 15
    L = lib()
    L.hpb_insitu_write_plasma.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int] + [C.c_double] * 5 + \
        [C.c_int, C.c_void_p]
    _check(L.hpb_insitu_write_plasma(str(path).encode(), time, step, a.shape[1], charge, mass, z_lo, z_hi,
                                     density_factor, int(normalized), a.ctypes.data), 'hpb_insitu_write_plasma')


def insitu_write_fields(path, time, step, z_lo, z_hi, normalized, dxdydz, sums):
    """hpb_insitu_write_fields (host only): raw per-slice sums[10, n_slices] -> one appended record"""
    a = np.ascontiguousarray(sums, dtype=np.float64)
    assert a.ndim == 2 and a.shape[0] == 10
    L = lib()
    L.hpb_insitu_write_fields.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                          C.c_int, C.c_double, C.c_void_p]
    _check(L.hpb_insitu_write_fields(str(path).encode(), time, step, a.shape[1], z_lo, z_hi, int(normalized),
                                     dxdydz, a.ctypes.data), 'hpb_insitu_write_fields')


def insitu_write_laser(path, time, step, z_lo, z_hi, normalized, dxdydz, nx, ny, sums):
    """hpb_insitu_write_laser (host only): raw per-slice values[8, n_slices] -> one appended record"""
    a = np.ascontiguousarray(sums, dtype=np.float64)
    assert a.ndim == 2 and a.shape[0] == 8
    L = lib()
    L.hpb_insitu_write_laser.argtypes = [C.c_char_p, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                         C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p]
    _check(L.hpb_insitu_write_laser(str(path).encode(), time, step, a.shape[1], z_lo, z_hi, int(normalized),
                                    dxdydz, nx, ny, a.ctypes.data), 'hpb_insitu_write_laser')


def read_insitu(path):
    """what tools/read_insitu_diagnostics.py of the reference does with one file: JSON dtype header,
    then raw records"""
    import json
    raw = open(path, 'rb').read()
    obj, off = json.JSONDecoder().raw_decode(raw.decode(errors='replace'))
    return np.frombuffer(raw, dtype=np.dtype(obj), offset=off)


def deck_check(deck: str, overrides: dict | None = None) -> dict:
    """Host-only dry run of the input-deck parser (no GPU): what hpb_sim_create would run, or
    HpbError with the same message (unsupported options abort like the reference's parser)."""
    L = lib()
    buf = C.create_string_buffer(8192)
    L.hpb_deck_check.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
    _check(L.hpb_deck_check(deck.encode(), _overrides_text(overrides), buf, len(buf)), 'hpb_deck_check')
    out = {}
    for kv in buf.value.decode().split(';'):
        if '=' in kv:
            k, v = kv.split('=', 1)
            try:
                out[k] = int(v)
            except ValueError:
                try:
                    out[k] = float(v)
                except ValueError:
                    out[k] = v
    return out


class Simulation:
    """HiPACE++ input deck -> slice loop on one GPU (host buffers at the boundary)."""

    def __init__(self, deck: str, overrides: dict | None = None, device: int = 0):
        self._h = C.c_void_p()
        self._L = lib()
        _check(self._L.hpb_sim_create(C.byref(self._h), deck.encode(), _overrides_text(overrides),
                                      device), 'hpb_sim_create')
        n = (C.c_int * 3)()
        lo = (C.c_double * 3)()
        hi = (C.c_double * 3)()
        self._L.hpb_sim_geometry(self._h, n, lo, hi)
        self.n_cell = tuple(n)
        self.prob_lo, self.prob_hi = tuple(lo), tuple(hi)
        self.nx, self.ny, self.nz = self.n_cell
        self.ncomp = self._L.hpb_sim_ncomp(self._h)
        self.ng = self._L.hpb_sim_nguard(self._h)      # guard cells (2 for the default order)

    def close(self):
        if self._h:
            self._L.hpb_sim_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: float):
        _check(self._L.hpb_sim_set_option(self._h, key.encode(), float(value)), 'hpb_sim_set_option')

    def evolve(self, step_begin: int = 0, step_end: int | None = None, n_slices: int = 0):
        """Hipace::Evolve for steps [step_begin, step_end] (default: up to the deck's max_step);
        returns the checksum dict."""
        _check(self._L.hpb_sim_evolve(self._h, step_begin, -1 if step_end is None else step_end, n_slices),
               'hpb_sim_evolve')
        return self.checksums()

    def begin_step(self, step: int = 0):
        _check(self._L.hpb_sim_begin_step(self._h, step), 'hpb_sim_begin_step')

    def solve_one_slice(self, islice: int):
        _check(self._L.hpb_sim_solve_one_slice(self._h, islice), 'hpb_sim_solve_one_slice')

    def comp_index(self, which_slice: str, name: str) -> int:
        return self._L.hpb_sim_comp_index(self._h, which_slice.encode(), name.encode())

    def field(self, name: str, which_slice: str = 'This') -> np.ndarray:
        """component incl. guard cells as a[j + g, i + g]"""
        c = self.comp_index(which_slice, name)
        if c < 0:
            raise KeyError((which_slice, name))
        out = np.empty((self.ny + 2 * self.ng, self.nx + 2 * self.ng))
        _check(self._L.hpb_sim_get_field(self._h, c, out.ctypes.data), 'hpb_sim_get_field')
        return out

    def set_field(self, name: str, arr, which_slice: str = 'This'):
        c = self.comp_index(which_slice, name)
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.shape == (self.ny + 2 * self.ng, self.nx + 2 * self.ng)
        _check(self._L.hpb_sim_set_field(self._h, c, a.ctypes.data), 'hpb_sim_set_field')

    def plasma_np(self, species: int = 0) -> int:
        return self._L.hpb_sim_plasma_np(self._h, species)

    def plasma(self, species: int = 0) -> dict:
        n = self.plasma_np(species)
        out = {}
        for k, nm in enumerate(PLASMA_REALS):
            a = np.empty(n)
            _check(self._L.hpb_sim_get_plasma_real(self._h, species, k, a.ctypes.data), 'get_plasma_real')
            out[nm] = a
        v = np.empty(n, dtype=np.uint8)
        _check(self._L.hpb_sim_get_plasma_valid(self._h, species, v.ctypes.data), 'get_plasma_valid')
        out['valid'] = v.astype(bool)
        return out

    def checksums(self) -> dict:
        n = self._L.hpb_sim_checksum_count(self._h)
        vals = np.empty(n)
        _check(self._L.hpb_sim_get_checksums(self._h, vals.ctypes.data), 'hpb_sim_get_checksums')
        return {self._L.hpb_sim_checksum_name(self._h, k).decode(): float(vals[k]) for k in range(n)}

    def beam_checksums(self, beam: int = 0) -> dict:
        out = (C.c_double * 9)()
        _check(self._L.hpb_sim_get_beam_checksums(self._h, beam, out), 'hpb_sim_get_beam_checksums')
        names = ('x', 'y', 'z', 'ux', 'uy', 'uz', 'w', 'id', 'count')
        return dict(zip(names, list(out)))

    def beam_np(self, beam: int = 0) -> int:
        return self._L.hpb_sim_beam_np(self._h, beam)

    def get_beam(self, host: dict, beam: int = 0):
        """D2H of the whole beam into host['real'] (7 x np float64), host['idcpu'] (np uint64),
        host['slot_off'] (nz+1 int64) -- preallocated (pinned) arrays."""
        ptrs = (C.c_void_p * 7)(*[host['real'][k].ctypes.data for k in range(7)])
        _check(self._L.hpb_sim_get_beam(self._h, beam, ptrs, host['idcpu'].ctypes.data,
                                        host['slot_off'].ctypes.data), 'hpb_sim_get_beam')

    def set_beam(self, host: dict, beam: int = 0):
        """H2D of a whole beam (enqueued on the simulation stream)."""
        ptrs = (C.c_void_p * 7)(*[host['real'][k].ctypes.data for k in range(7)])
        _check(self._L.hpb_sim_set_beam(self._h, beam, ptrs, host['idcpu'].ctypes.data,
                                        host['slot_off'].ctypes.data), 'hpb_sim_set_beam')

    def beam_slice_capacity(self, beam: int = 0) -> int:
        return self._L.hpb_sim_beam_slice_capacity(self._h, beam)

    # -- time-step pipeline (MultiBuffer's role, NCCL point-to-point) ---------------------------
    def pipeline_init(self, rank: int, world: int, dist=None):
        """Join the ring of `world` GPUs: this rank owns the time steps rank, rank + world, ...
        The ncclUniqueId of every edge r -> r+1 is created by rank r and exchanged through
        torch.distributed (any backend; `dist` is the initialised torch.distributed module)."""
        from . import pipeline as pl
        self.rank, self.world = rank, world
        if world == 1:
            _check(self._L.hpb_sim_pipeline_init(self._h, 0, 1, None, None), 'hpb_sim_pipeline_init')
            return
        my_id = C.create_string_buffer(NCCL_ID_BYTES)
        _check(self._L.hpb_nccl_unique_id(my_id), 'hpb_nccl_unique_id')
        ids = pl.exchange_edge_ids(dist, rank, world, my_id.raw)
        _check(self._L.hpb_sim_pipeline_init(self._h, rank, world, ids[pl.upstream(rank, world)],
                                             ids[rank]), 'hpb_sim_pipeline_init')

    def timer_start(self):
        _check(self._L.hpb_sim_timer_start(self._h), 'hpb_sim_timer_start')

    def timer_stop(self) -> float:
        ms = C.c_double(0.)
        _check(self._L.hpb_sim_timer_stop(self._h, C.byref(ms)), 'hpb_sim_timer_stop')
        return ms.value

    def beam_packet(self, slot: int, beam: int = 0) -> np.ndarray:
        """wire message of one slice packet of the current beam ring (uint8)"""
        n = 64 + 64 * self.beam_slice_capacity(beam)
        out = np.empty(n, dtype=np.uint8)
        _check(self._L.hpb_sim_get_beam_packet(self._h, beam, slot, out.ctypes.data), 'hpb_sim_get_beam_packet')
        return out

    def pipeline_message_bytes(self) -> int:
        return self._L.hpb_sim_pipeline_message_bytes(self._h)

    def run(self, max_step: int, rank: int = 0, world: int = 1):
        """Hipace::Evolve on this rank: the time steps rank, rank + world, ... <= max_step
        (Hipace.cpp:401).  Returns the summed stats of the owned steps."""
        from . import pipeline as pl
        self.set_option('max_step', max_step)
        tot = {}
        for step in pl.owned_steps(rank, world, max_step):
            self.evolve(step, step)
            for k, v in self.stats().items():
                tot[k] = tot.get(k, 0) + v
        return tot

    def mg_iters(self) -> list:
        """V-cycles of every slice of the last evolve (head first)"""
        n = self._L.hpb_sim_get_mg_iters(self._h, None, 0)
        buf = (C.c_int * max(n, 1))()
        self._L.hpb_sim_get_mg_iters(self._h, buf, n)
        return list(buf[:n])

    def time(self):
        """(physical time, dt) of the step that ran last"""
        t, dt = C.c_double(0.), C.c_double(0.)
        _check(self._L.hpb_sim_get_time(self._h, C.byref(t), C.byref(dt)), 'hpb_sim_get_time')
        return t.value, dt.value

    def stats(self) -> dict:
        st = hpb_sim_stats()
        _check(self._L.hpb_sim_get_stats(self._h, C.byref(st)), 'hpb_sim_get_stats')
        return {f: getattr(st, f) for f, _ in st._fields_}


def set_global_option(key: str, value: float):
    """process-wide library switches (hpb_set_option with a NULL context): "pdl", "bluestein_min_prime" """
    _check(lib().hpb_set_option(None, key.encode(), C.c_double(value)), 'hpb_set_option')


def measure_fp64_peak(device: int = 0, reps: int = 5) -> float:
    """fp64 FMA peak of the device in TFLOP/s (measured, csrc/peaks.cu)"""
    out = C.c_double(0.)
    _check(lib().hpb_measure_fp64_peak(device, reps, C.byref(out)), 'hpb_measure_fp64_peak')
    return out.value


class Context:
    """Kernel seams over caller-owned device memory (torch CUDA tensors)."""

    def __init__(self, nx, ny, dx, dy, dz, x_off, y_off, *, normalized=True, c=1., ep0=1., mu0=1.,
                 q_e=1., m_e=1., stream=None):
        self._L = lib()
        self.geom = hpb_geom(nx, ny, dx, dy, dz, x_off, y_off, c, ep0, mu0, q_e, m_e,
                             1 if normalized else 0)
        self._h = C.c_void_p()
        _check(self._L.hpb_create(C.byref(self._h), C.byref(self.geom), stream), 'hpb_create')
        self.nx, self.ny = nx, ny

    def close(self):
        if self._h:
            self._L.hpb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- views over torch tensors ----------------------------------------------------------
    def slice_view(self, t) -> hpb_slice:
        """t: float64 CUDA tensor [ncomp, ny + 2g, nx + 2g] (component-major slice array); g from
        the tensor's shape (2 unless set_deposition_order chose another order)"""
        ncomp, ny_t, nx_t = t.shape
        assert t.is_cuda and t.is_contiguous() and str(t.dtype) == 'torch.float64'
        g = (nx_t - self.nx) // 2
        assert nx_t == self.nx + 2 * g and ny_t == self.ny + 2 * g
        return hpb_slice(t.data_ptr(), -g, -g, nx_t, ny_t, nx_t, nx_t * ny_t, ncomp)

    def set_deposition_order(self, order_xy: int, derivative_type: int = 2):
        """hipace.depos_order_xy / hipace.depos_derivative_type for the particle kernels"""
        _check(self._L.hpb_set_deposition_order(self._h, order_xy, derivative_type),
               'hpb_set_deposition_order')

    @staticmethod
    def plasma_view(reals, idcpu) -> hpb_plasma:
        """reals: float64 CUDA tensor [11, np] (PlasmaIdx order); idcpu: int64 tensor [np]"""
        p = hpb_plasma()
        n = reals.shape[1]
        for k in range(11):
            p.r[k] = reals[k].data_ptr()
        p.idcpu = idcpu.data_ptr()
        p.np = n
        return p

    @staticmethod
    def comps(**kw):
        arr = (C.c_int * len(COMP_IDS))(*([-1] * len(COMP_IDS)))
        for k, v in kw.items():
            arr[COMP[k.upper()]] = v
        return arr

    # -- seams -----------------------------------------------------------------------------
    def deposit_current(self, pl, sl, charge, mass, *, jx=-1, jy=-1, rho=-1, chi=-1, rhomjz=-1,
                        max_qsa=35., n_qsa_ptr=None):
        _check(self._L.hpb_deposit_current(self._h, pl, sl, charge, mass, jx, jy, rho, chi, rhomjz,
                                           max_qsa, n_qsa_ptr), 'hpb_deposit_current')

    def explicit_deposition(self, pl, sl, charge, mass, comps):
        _check(self._L.hpb_explicit_deposition(self._h, pl, sl, charge, mass, comps),
               'hpb_explicit_deposition')

    def advance_plasma_particles(self, pl, sl, charge, mass, comps, *, n_subcycles=1,
                                 temp_slice=False, bc='Periodic', bc_lo=(0., 0.), bc_hi=(0., 0.)):
        lo = (C.c_double * 2)(*bc_lo)
        hi = (C.c_double * 2)(*bc_hi)
        _check(self._L.hpb_advance_plasma_particles(self._h, pl, sl, charge, mass, n_subcycles,
                                                    int(temp_slice), BC[bc], lo, hi, comps),
               'hpb_advance_plasma_particles')

    @staticmethod
    def beam_view(reals, idcpu, counts=None, cap=None) -> hpb_beam_slice:
        """reals: float64 CUDA tensor [7, cap] (x y z w ux uy uz); idcpu: int64 tensor [cap];
        counts: int64 CUDA tensor [2] = (np, np incl. slipped) or None"""
        b = hpb_beam_slice()
        for k, nm in enumerate(('x', 'y', 'z', 'w', 'ux', 'uy', 'uz')):
            setattr(b, nm, reals[k].data_ptr())
        b.idcpu = idcpu.data_ptr()
        b.np = reals.shape[1] if cap is None else cap
        b.d_np = counts.data_ptr() if counts is not None else None
        return b

    def extfields(self, exprs):
        """six expressions Ex Ey Ez Bx By Bz of (x, y, z, t) -> handle"""
        arr = (C.c_char_p * 6)(*[e.encode() for e in exprs])
        h = C.c_void_p()
        _check(self._L.hpb_extfields_create(C.byref(h), arr), 'hpb_extfields_create')
        return h

    def advance_beam_particles(self, bm, nsub, sl, charge, mass, comps, *, n_subcycles=10, dt=0.,
                               time=0., min_z=-1e300, do_z_push=True, bc='Periodic',
                               bc_lo=(0., 0.), bc_hi=(0., 0.), ext=None, class_counts=None,
                               checksum=None):
        lo = (C.c_double * 2)(*bc_lo)
        hi = (C.c_double * 2)(*bc_hi)
        _check(self._L.hpb_advance_beam_particles(
            self._h, bm, nsub.data_ptr(), sl, charge, mass, n_subcycles, dt, time, min_z,
            int(do_z_push), BC[bc], lo, hi, comps, ext,
            class_counts.data_ptr() if class_counts is not None else None,
            checksum.data_ptr() if checksum is not None else None), 'hpb_advance_beam_particles')

    def beam_shift_slipped(self, bm, nsub, min_z, class_counts, stay, stay_counts, nxt, next_counts,
                           next_nsub, overflow):
        _check(self._L.hpb_beam_shift_slipped(
            self._h, bm, nsub.data_ptr(), min_z, class_counts.data_ptr(), stay,
            stay_counts.data_ptr(), nxt, next_counts.data_ptr() if next_counts is not None else None,
            next_nsub.data_ptr() if next_nsub is not None else None, overflow.data_ptr()),
            'hpb_beam_shift_slipped')

    def poisson_solve(self, rhs, sl, c_lhs):
        """rhs: float64 CUDA tensor [nbatch, ny, nx]; c_lhs: destination components"""
        nb = rhs.shape[0]
        arr = (C.c_int * nb)(*c_lhs)
        _check(self._L.hpb_poisson_solve(self._h, rhs.data_ptr(), sl, arr, nb), 'hpb_poisson_solve')

    def poisson_solve_periodic(self, rhs, sl, c_lhs):
        """fields.poisson_solver = FFTPeriodic; arguments as poisson_solve"""
        nb = rhs.shape[0]
        arr = (C.c_int * nb)(*c_lhs)
        _check(self._L.hpb_poisson_solve_periodic(self._h, rhs.data_ptr(), sl, arr, nb),
               'hpb_poisson_solve_periodic')

    def enforce_periodic(self, sl, do_sum, comp_list):
        """Fields::EnforcePeriodic: SumBoundary (do_sum) or FillBoundary of the listed components"""
        arr = (C.c_int * len(comp_list))(*comp_list)
        _check(self._L.hpb_fields_enforce_periodic(self._h, sl, int(bool(do_sum)), arr, len(comp_list)),
               'hpb_fields_enforce_periodic')

    def solve_psi_ez_bz(self, sl, comps):
        _check(self._L.hpb_fields_solve_psi_ez_bz(self._h, sl, comps), 'hpb_fields_solve_psi_ez_bz')

    def sxsy_from_beam(self, sl, comps):
        _check(self._L.hpb_fields_sxsy_from_beam(self._h, sl, comps), 'hpb_fields_sxsy_from_beam')

    def mg_solve2(self, t_sol, t_rhs, t_acf_r, acf_i, tol_rel=1e-4, tol_abs=0.0, max_iters=200) -> int:
        """hpmg solve2 on float64 CUDA tensors: t_sol [2, ny, nx] (initial guess in, solution out),
        t_rhs [2, ny, nx], t_acf_r [ny, nx]; acf_i scalar"""
        it = C.c_int(0)
        _check(self._L.hpb_mg_solve2(self._h, C.c_void_p(t_sol.data_ptr()), C.c_void_p(t_rhs.data_ptr()),
                                     C.c_void_p(t_acf_r.data_ptr()), C.c_double(acf_i), C.c_double(tol_rel),
                                     C.c_double(tol_abs), max_iters, C.byref(it)), 'hpb_mg_solve2')
        return it.value

    def mg_solve1(self, sl, c_sol, c_rhs, c_acf, tol_rel=1e-4, tol_abs=np.finfo(float).tiny,
                  max_iters=200) -> int:
        it = C.c_int(0)
        _check(self._L.hpb_mg_solve1(self._h, sl, c_sol, c_rhs, c_acf, tol_rel, tol_abs, max_iters,
                                     C.byref(it)), 'hpb_mg_solve1')
        return it.value
