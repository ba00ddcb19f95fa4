"""GPU parity of boundary.field = Periodic (Fields::EnforcePeriodic, fields/Fields.cpp:1117-1145, called at
:859-861, :920-922 and Hipace.cpp:817-821, :924-927) and fields.poisson_solver = FFTPeriodic
(FFTPoissonSolverPeriodic.cpp:69-149) against the oracle.  No reference golden exists for periodic fields
(no deck of the reference uses them): parity unpinned, CUDA == oracle is what is shown here."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL_SUM = 1e-9
RTOL_CELL = 1e-9
FLOOR = 1e-3


@pytest.mark.parametrize('nx,ny', [(64, 64), (48, 40), (63, 45), (250, 128), (1024, 1024)])
def test_periodic_poisson_seam_matches_oracle(nx, ny):
    """hpb_poisson_solve_periodic on caller memory: 1, 2 and 3 right-hand sides (two ride in the Re / Im
    lanes of one complex transform), incl. odd and prime-factor-rich lengths"""
    import torch
    import hipace_b200 as hp
    from oracle.hipace_oracle import poisson_periodic
    rng = np.random.default_rng(nx * 1000 + ny)
    dx, dy = 16. / nx, 12. / ny
    ctx = hp.Context(nx, ny, dx, dy, 0.1, -8 + dx / 2, -6 + dy / 2)
    g = hp.NGUARD
    for nb in (3, 1, 2):
        rhs = rng.standard_normal((nb, ny, nx))
        sl_t = torch.full((4, ny + 2 * g, nx + 2 * g), 7.0, dtype=torch.float64, device='cuda')
        comps = [0, 2, 3][:nb]
        ctx.poisson_solve_periodic(torch.from_numpy(rhs).cuda(), ctx.slice_view(sl_t), comps)
        torch.cuda.synchronize()
        out = sl_t.cpu().numpy()
        for b, c in enumerate(comps):
            want = poisson_periodic(rhs[b], dx, dy)
            err = np.abs(out[c, g:-g, g:-g] - want).max() / np.abs(want).max()
            assert err <= 1e-11, (nb, b, err)
        assert (out[1] == 7.0).all()
        assert (out[0, :g] == 7.0).all() and (out[0, :, :g] == 7.0).all()
    ctx.close()


@pytest.mark.parametrize('nx,ny,order', [(64, 64, 2), (37, 51, 2), (40, 34, 0), (48, 48, 3)])
def test_enforce_periodic_seam_matches_oracle(nx, ny, order):
    """hpb_fields_enforce_periodic on caller memory for every guard width (orders 0, 2, 3 -> g = 1, 2, 3):
    SumBoundary bit-exact (same association as the oracle: cell + x image + y image + corner image),
    FillBoundary bit-exact, untouched components untouched"""
    import torch
    import hipace_b200 as hp
    from oracle.hipace_oracle import enforce_periodic
    rng = np.random.default_rng(nx + ny + order)
    g = (order + 1) // 2 + 1
    ctx = hp.Context(nx, ny, 0.1, 0.1, 0.1, 0., 0.)
    ctx.set_deposition_order(order, 2)
    a = rng.standard_normal((5, ny + 2 * g, nx + 2 * g))
    for do_sum in (True, False):
        t = torch.from_numpy(a.copy()).cuda()
        ctx.enforce_periodic(ctx.slice_view(t), do_sum, [0, 3, -1, 4])
        torch.cuda.synchronize()
        got = t.cpu().numpy()
        want = a.copy()
        enforce_periodic([want[0], want[3], want[4]], g, do_sum)
        assert np.array_equal(got, want), do_sum
    ctx.close()


@pytest.mark.parametrize('fuse', [0, 1])
@pytest.mark.parametrize('ov', [
    {'boundary.field': 'Periodic', 'fields.poisson_solver': 'FFTPeriodic'},
    {'boundary.field': 'Periodic'},
    {'fields.poisson_solver': 'FFTPeriodic'},
], ids=['periodic+FFTPeriodic', 'periodic+Dirichlet_solver', 'FFTPeriodic_only'])
def test_periodic_slice_loop_matches_oracle(ov, fuse, repo_root):
    """the blow-out wake in a box narrow enough for the sheath and the return current to reach the edges
    (the periodic images matter), 60 slices, both driver orders: field checksums, the solved fields per
    cell INCLUDING their guard cells (FillBoundary), V-cycle counts, the full particle state"""
    import hipace_b200 as hp
    from oracle.hipace_oracle import Simulation as Oracle
    text = open(os.path.join(repo_root, 'examples', 'blowout_wake_normalized.in')).read()
    ov = dict(ov, **{'amr.n_cell': '64 48 100', 'geometry.prob_lo': '-5. -4. -6.', 'geometry.prob_hi': '5. 4. 6.'})
    nsl = 60
    ref = Oracle(text, ov)
    want = ref.evolve(nsl)
    sim = hp.Simulation(text, ov)
    sim.set_option('fuse', fuse)
    got = sim.evolve(0, 0, nsl)
    atol = 1e-12 * max(abs(w) for w in want.values())
    for k, w in want.items():
        assert abs(got[k] - w) <= RTOL_SUM * abs(w) + atol, (k, got[k], w)
    assert [g_ for g_, w in zip(sim.mg_iters(), ref.mg_cycles) if w > 0] == [w for w in ref.mg_cycles if w > 0]
    assert max(ref.mg_cycles) >= 2
    for n in ('Psi', 'Ez', 'Bz', 'Bx', 'By', 'ExmBy', 'EypBx'):
        a, b = sim.field(n), ref.T(n)
        err = np.abs(a - b).max() / max(np.abs(b).max(), FLOOR)
        assert err <= RTOL_CELL, (n, err)
    p, o = sim.plasma(), ref.plasmas[0]
    assert np.array_equal(p['valid'], o.valid)
    v = o.valid
    for nm, ov_ in (('x', o.x), ('y', o.y), ('ux', o.ux), ('uy', o.uy), ('psi', o.psi), ('w', o.w)):
        scale = max(np.abs(ov_[v]).max(), 1e-300)
        assert np.abs(p[nm][v] - ov_[v]).max() / scale <= 1e-9, nm
    # the periodic images did matter in this box
    if 'boundary.field' in ov:
        d = Oracle(text, dict(ov, **{'boundary.field': 'Dirichlet'})).evolve(nsl)
        assert any(abs(d[k] - w) > 1e-6 * abs(w) for k, w in want.items() if w)
    sim.close()
