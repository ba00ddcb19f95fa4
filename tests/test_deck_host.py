"""Host logic without a GPU: the C++ input-deck parser behind hpb_sim_create (ParmParse subset +
HiPACE++ expression parser, src/utils/Parser.H:316-395) against the oracle's reading of the same
decks, and the 'unsupported' errors that stand in for the reference's aborts."""
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# decks the oracle runs but the CUDA driver must refuse (with the reason), not mis-run: none any more
# (the laser envelope advance with the multigrid solver, hpmg type 2, was the last one)
ORACLE_ONLY = {}
DECKS = sorted(p for p in glob.glob(os.path.join(ROOT, 'examples', '*.in'))
               if os.path.basename(p) not in ORACLE_ONLY)


@pytest.mark.parametrize('path', DECKS, ids=[os.path.basename(p) for p in DECKS])
def test_parser_agrees_with_oracle(path):
    import hipace_b200 as hp
    from oracle.hipace_oracle import Simulation as Oracle
    text = open(path).read()
    got = hp.deck_check(text)
    ref = Oracle(text, {})
    g = ref.geom
    assert (got['nx'], got['ny'], got['nz']) == (g.nx, g.ny, g.nz)
    for k, w in (('dx', g.dx), ('dy', g.dy), ('dz', g.dz)):
        assert got[k] == pytest.approx(w, rel=1e-15)
    assert got['n_plasmas'] == len(ref.plasmas) and got['n_beams'] == len(ref.beams)
    for k, p in enumerate(ref.plasmas):
        assert got[f'plasma{k}.charge'] == pytest.approx(p.charge, rel=1e-15)
        assert got[f'plasma{k}.mass'] == pytest.approx(p.mass, rel=1e-15)
        assert got[f'plasma{k}.neutralize'] == int(p.neutralize_background)
        # (constants defined through other constants with blanks in the expression, SI deck)
        assert got[f'plasma{k}.density0'] == pytest.approx(float(p.density(0., 0., 0.)), rel=1e-14)
    for k, b in enumerate(ref.beams):
        assert got[f'beam{k}.charge'] == pytest.approx(b.charge, rel=1e-15)
        assert got[f'beam{k}.mass'] == pytest.approx(b.mass, rel=1e-15)


def test_deposition_orders_and_the_combination_the_reference_rejects():
    import hipace_b200 as hp
    text = open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()
    for order in range(4):
        for dtype in range(3):
            ov = {'hipace.depos_order_xy': order, 'hipace.depos_derivative_type': dtype}
            if (order, dtype) == (0, 0):          # Hipace.cpp:52-53
                with pytest.raises(hp.HpbError) as e:
                    hp.deck_check(text, ov)
                assert 'would vanish' in str(e.value)
            else:
                hp.deck_check(text, ov)
    for bad in ({'hipace.depos_order_xy': 4}, {'hipace.depos_derivative_type': 3}):
        with pytest.raises(hp.HpbError):
            hp.deck_check(text, bad)


def test_overrides_constants_and_expressions():
    import hipace_b200 as hp
    text = open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()
    got = hp.deck_check(text, {'amr.n_cell': '1024 1023 7', 'plasma.ppc': '3 2', 'hipace.dt': '2.5*2',
                               'my_constants.kp_inv': 2., 'geometry.prob_hi': '4*kp_inv 8. 6.'})
    assert (got['nx'], got['ny'], got['nz']) == (1024, 1023, 7)
    assert got['plasma0.ppc'] == '3x2' and got['dt'] == 5.0
    assert got['dx'] == pytest.approx(16. / 1024, rel=1e-15)
    # GetPosOffset with the grown box reduces to prob_lo + dx/2 (src/fields/Fields.H:71-77)
    assert got['x_off'] == pytest.approx(-8. + 0.5 * got['dx'], rel=1e-14)


@pytest.mark.parametrize('ov,msg', [
    ({'beam.injection_type': 'fixed_weight'}, 'fixed_ppc'),
    ({'hipace.depos_order_xy': 4}, 'depos_order_xy'),
    ({'hipace.bxby_solver': 'semi-implicit'}, 'bxby_solver'),
    ({'boundary.field': 'Open'}, 'predictor-corrector'),
    ({'boundary.field': 'Periodic', 'hipace.bxby_solver': 'predictor-corrector'}, 'explicit'),
    ({'fields.poisson_solver': 'MGDirichlet'}, 'MGDirichlet'),
    ({'boundary.field': 'Mirror'}, 'Dirichlet, Periodic or Open'),
    ({'plasma.u_std': '0. 0. 1e-3'}, 'RNG'),
    ({'amr.n_cell': '64 64'}, '3 values'),
    ({'diagnostic.diag_type': 'yz'}, 'diag_type'),
    ({'beam.profile': 'parabolic'}, 'profile'),
    # physics-changing options of the reference this implementation does not have: refused, never
    # silently ignored (unknown keys used to be dropped)
    ({'plasma.can_ionize': 1}, 'plasma.can_ionize'),
    ({'plasma.ionization_product': 'elec'}, 'plasma.ionization_product'),
    ({'hipace.collisions': 'c1'}, 'hipace.collisions'),
    ({'plasma.do_symmetrize': 1}, 'plasma.do_symmetrize'),
    ({'plasma.temperature_in_ev': 10.}, 'plasma.temperature_in_ev'),
    ({'plasma.fine_patch(x,y)': '1'}, 'fine_patch'),
    ({'beam.do_salame': 1}, 'beam.do_salame'),
    ({'beam.random_ppc': '1 1 1'}, 'beam.random_ppc'),
    ({'hipace.max_time': 3.}, 'max_time'),
    ({'amr.max_level': 1}, 'max_level'),
    ({'hipace.do_beam_jz_minus_rho': 1}, 'do_beam_jz_minus_rho'),
    ({'hipace.no_such_option': 1}, 'hipace.no_such_option'),
])
def test_unsupported_options_fail_loudly(ov, msg):
    import hipace_b200 as hp
    text = open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()
    with pytest.raises(hp.HpbError) as e:
        hp.deck_check(text, ov)
    assert msg in str(e.value)


def test_inert_and_output_only_keys_are_accepted():
    """keys that cannot change the physics of the path (verbosity, diagnostics selection, AMReX
    runtime switches) and blocks of names the deck does not activate are accepted, like the reference"""
    import hipace_b200 as hp
    text = open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()
    hp.deck_check(text, {'hipace.verbose': 2, 'diagnostic.output_period': 3, 'amrex.the_arena_is_managed': 1,
                         'amr.max_level': 0, 'other_beam.density': 3.})


def test_laser_deck_is_accepted_with_both_envelope_solvers():
    import hipace_b200 as hp
    import json
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'laser_blowout_wake_explicit.SI.1Rank.json')))
    text = open(os.path.join(ROOT, meta['deck'])).read()
    got = hp.deck_check(text, meta['overrides'])
    assert got['n_lasers'] == 1 and got['n_beams'] == 0 and (got['nx'], got['nz']) == (128, 100)
    # the envelope advance over time steps: the reference's default (multigrid, hpmg type 2) and fft
    for ov in ({'max_step': 2}, {'max_step': 2, 'lasers.solver_type': 'fft'},
               {'max_step': 2, 'lasers.MG_tolerance_rel': 1e-5, 'lasers.MG_average_rhs': 0}):
        hp.deck_check(text, dict(meta['overrides'], **ov))
    with pytest.raises(hp.HpbError) as e:
        hp.deck_check(text, dict(meta['overrides'], **{'lasers.solver_type': 'spectral'}))
    assert 'solver_type' in str(e.value)


def test_no_gpu_means_error_not_fallback():
    """the product path must fail loudly without a device (no CPU fallback)"""
    import torch
    import hipace_b200 as hp
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    text = open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()
    with pytest.raises(hp.HpbError) as e:
        hp.Simulation(text, {})
    assert 'no CUDA device' in str(e.value)


def test_command_line_driver_check_mode():
    """hpb200_run (C++ host over the C-ABI, the reference's `hipace inputs key=value ...`)"""
    import subprocess
    from hipace_b200.build import build_library
    build_library()
    exe = os.path.join(ROOT, 'hipace_b200', 'bin', 'hpb200_run')
    deck = os.path.join(ROOT, 'examples', 'linear_wake_normalized.in')
    p = subprocess.run([exe, '--check', deck, 'amr.n_cell=64 64 64', 'max_step=3'], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert 'nx=64;ny=64;nz=64;' in p.stdout and 'max_step=3;' in p.stdout
    p = subprocess.run([exe, '--check', deck, 'hipace.depos_order_xy=5'], capture_output=True, text=True)
    assert p.returncode == 1 and 'depos_order_xy' in p.stderr


@pytest.mark.parametrize('n,ppc', [(1024 * 1024, 4), (63 * 63, 4), (100 * 36, 9), (5, 2), (128, 1)])
def test_push_thread_maps_are_permutations(n, ppc):
    """the lattice-ordered thread -> particle maps of the push kernel (linear, passes interleaved
    warp by warp, passes interleaved CTA by CTA) each visit every particle exactly once"""
    import ctypes as C
    import numpy as np
    import hipace_b200 as hp
    L = hp.lib()
    L.hpb_debug_push_thread_map.restype = C.c_long
    L.hpb_debug_push_thread_map.argtypes = [C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_long]
    for mode in (0, 1, 2):
        nw = L.hpb_debug_push_thread_map(n, ppc, mode, None, 0)
        out = np.full(nw * 32, -2, dtype=np.int64)
        assert L.hpb_debug_push_thread_map(n, ppc, mode, out.ctypes.data, out.size) == nw
        ids = out[out >= 0]
        assert ids.size == n * ppc and np.array_equal(np.sort(ids), np.arange(n * ppc)), mode
        assert (out >= -1).all()
        if mode == 2 and ppc > 1 and n >= 256:
            # the four warps of a CTA work on 128 consecutive cells of ONE pass ...
            cta0 = out[:128]
            assert np.array_equal(cta0, np.arange(128))
            # ... and the next CTA on the same cells of the next pass
            assert np.array_equal(out[128:256], n + np.arange(128))


@pytest.mark.parametrize('p', [7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61])
def test_prime_stage_tensor_core_table(p):
    """the fragment-ordered coefficient table of the odd-prime FFT stage (fft_smem.cuh): with it, the padded
    8 x 4 tiles of mma.m8n8k4 (lane l holds A[l / 4][l % 4]) must reproduce the DFT of length p through
    P = C U, Q = S V, X_b = x_0 + P_b - i Q_b, X_{p-b} = x_0 + P_b + i Q_b -- emulated here tile by tile in
    NumPy exactly as the kernel walks the table"""
    import ctypes as C
    import numpy as np
    import hipace_b200 as hp
    L = hp.lib()
    L.hpb_debug_fft_prime_table.restype = C.c_long
    L.hpb_debug_fft_prime_table.argtypes = [C.c_int, C.c_void_p, C.c_long]
    n = L.hpb_debug_fft_prime_table(p, None, 0)
    h = (p - 1) // 2
    KT, MT = (h + 3) // 4, (h + 8) // 8
    assert n == MT * KT * 64
    tab = np.zeros(n)
    assert L.hpb_debug_fft_prime_table(p, tab.ctypes.data, n) == n
    tab = tab.reshape(MT, KT, 2, 32)
    rng = np.random.default_rng(p)
    x = rng.standard_normal(p) + 1j * rng.standard_normal(p)
    U = np.array([x[t] + x[p - t] for t in range(1, h + 1)])          # rows t = 1 .. h
    V = np.array([x[t] - x[p - t] for t in range(1, h + 1)])
    X = np.zeros(p, complex)
    lanes = np.arange(32)
    for mt in range(MT):
        Pm = np.zeros(8, complex)
        Qm = np.zeros(8, complex)
        for kt in range(KT):
            A_c = np.zeros((8, 4)); A_s = np.zeros((8, 4))
            A_c[lanes // 4, lanes % 4] = tab[mt, kt, 0]
            A_s[lanes // 4, lanes % 4] = tab[mt, kt, 1]
            t = 4 * kt + np.arange(4)
            ok = t < h
            bu = np.where(ok, U[np.minimum(t, h - 1)], 0.)
            bv = np.where(ok, V[np.minimum(t, h - 1)], 0.)
            Pm += A_c @ bu
            Qm += A_s @ bv
        for r in range(8):
            b = 8 * mt + r
            if b > h:
                assert abs(Pm[r]) == 0. and abs(Qm[r]) == 0.          # padded rows are exact zeros
                continue
            X[b] = x[0] + Pm[r] - 1j * Qm[r]
            if b:
                X[p - b] = x[0] + Pm[r] + 1j * Qm[r]
    want = np.fft.fft(x)
    assert np.abs(X - want).max() <= 1e-13 * np.abs(want).max()
