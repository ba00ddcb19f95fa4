"""GPU parity of the beam path: AdvanceBeamParticlesSlice, external fields, shiftSlippedParticles
and the slice-packet ring that hands a beam from one time step to the next, through the C-ABI.

  * the reference's own golden beam_evolution.1Rank (21 time steps, dt = 3, external focusing
    field, tests/beam_evolution.1Rank.sh) -- pins push + external fields + multi-step hand-off;
  * a slow beam that slips backwards through the slices, against the oracle: per-slice particle
    counts and particle ids (order included) bit-exact, phase space to 1e-9;
  * the kernel seams on caller-owned memory (host counts, d_np == NULL).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
RTOL_SUM = 1e-9


def _deck(repo_root, name):
    return open(os.path.join(repo_root, 'examples', name)).read()


def test_beam_evolution_matches_reference_golden(repo_root):
    import hipace_b200 as hp
    meta = json.load(open(os.path.join(GOLD, 'beam_evolution.1Rank.json')))
    sim = hp.Simulation(open(os.path.join(repo_root, meta['deck'])).read(), meta['overrides'])
    nlast = int(meta['overrides']['max_step'])
    sim.evolve(0, nlast - 1)
    # theory (examples/beam_in_vacuum/analysis_beam_push.py): sigma_x = sigma_0 |cos(omega_beta t)|
    # at the time of the last output, t = max_step * dt (the beam has been pushed max_step times)
    host = _host_beam(sim)
    sim.get_beam(host)
    x, w = host['real'][0], host['real'][3]
    t = nlast * float(meta['overrides']['hipace.dt'])
    std_theory = 0.5 * abs(np.cos(np.sqrt(0.5 / 1000.) * t))
    std_sim = np.sqrt((x * x * w).sum() / w.sum())
    assert abs(std_sim - std_theory) / std_theory < 2e-3
    cs = sim.evolve(nlast, nlast)
    gold = meta['checksums']
    for name, want in gold['lev=0'].items():
        assert abs(cs[name] - want) <= RTOL_SUM * abs(want) + 1e-40, (name, cs[name], want)
    bc = sim.beam_checksums()
    for name, want in gold['beam'].items():
        if name in bc:
            assert abs(bc[name] - want) <= RTOL_SUM * abs(want) + 1e-40, (name, bc[name], want)
    assert bc['count'] == gold['beam']['charge']          # |q| = 1: number of particles, bit-exact
    sim.close()


def _host_beam(sim, beam=0):
    n = max(sim.beam_np(beam), 1)
    return {'real': np.zeros((7, n)), 'idcpu': np.zeros(n, dtype=np.uint64),
            'slot_off': np.zeros(sim.nz + 1, dtype=np.int64)}


SLIP_OV = {'amr.n_cell': '48 48 20', 'hipace.dt': 4., 'beam.u_mean': '0. 0. 3.', 'beam.ppc': '1 1 2',
           'beam.zmin': -3., 'beam.zmax': 3., 'beam.density': 0.5, 'beam.n_subcycles': 4,
           'beam.radius': 3.}


@pytest.mark.parametrize('bc', ['Periodic'])
def test_slipping_beam_matches_oracle(bc, repo_root):
    """gamma ~ 3 beam: dz/dt = beta - 1 = -0.05, so with dt = 4 a third of a cell per step slips
    into the next slice; 4 time steps.  Counts / ids per slice bit-exact against the oracle."""
    import hipace_b200 as hp
    from oracle.hipace_oracle import Simulation as Oracle
    from hipace_b200.pipeline import unpack_idcpu
    text = _deck(repo_root, 'blowout_wake_normalized.in')
    ov = dict(SLIP_OV, **{'boundary.particle': bc})
    nsteps = 4
    ref = Oracle(text, ov)
    ref.evolve(step_end=nsteps - 1)
    sim = hp.Simulation(text, ov)
    cs = sim.evolve(0, nsteps - 1)
    for name, want in ref.checksums.items():
        assert abs(cs[name] - want) <= 1e-8 * abs(want) + 1e-30, (name, cs[name], want)
    host = _host_beam(sim)
    sim.get_beam(host)
    off = host['slot_off']
    ids, valid = unpack_idcpu(host['idcpu'])
    b = ref.beams[0]
    nz = ref.geom.nz
    n_moved = 0
    for isl in range(nz):
        slot = nz - 1 - isl
        o = b.slices[isl]
        n = o['np']
        assert off[slot + 1] - off[slot] == n, (isl, off[slot + 1] - off[slot], n)
        sl = slice(off[slot], off[slot + 1])
        assert np.array_equal(ids[sl], o['id'][:n]), isl                     # same particles, same order
        assert valid[sl].all()
        for k, nm in enumerate(('x', 'y', 'z', 'w', 'ux', 'uy', 'uz')):
            want = o[nm][:n]
            if n:
                err = np.abs(host['real'][k][sl] - want).max() / max(np.abs(want).max(), 1e-300)
                assert err <= 1e-9, (isl, nm, err)
        n_moved += int(n)
    assert n_moved == off[nz]
    # the test only means something if particles actually changed slice
    first = Oracle(text, ov)
    for isl in range(nz):
        first.beam_slice(first.beams[0], isl)
    changed = sum(first.beams[0].slices[i]['np'] != b.slices[i]['np'] for i in range(nz))
    assert changed >= 3, changed
    st = sim.stats()
    assert st['n_beam_pushed'] == ref.n_beam_pushed
    sim.close()


def test_set_get_beam_round_trip(repo_root):
    import hipace_b200 as hp
    text = _deck(repo_root, 'blowout_wake_normalized.in')
    sim = hp.Simulation(text, {'amr.n_cell': '32 32 24'})
    host = _host_beam(sim)
    sim.get_beam(host)
    n = host['slot_off'][-1]
    assert n == sim.beam_np() and n > 0
    rng = np.random.default_rng(3)
    new = {'real': rng.standard_normal((7, n)), 'idcpu': host['idcpu'].copy(),
           'slot_off': host['slot_off'].copy()}
    sim.set_beam(new)
    back = _host_beam(sim)
    sim.get_beam(back)
    assert np.array_equal(back['slot_off'], new['slot_off'])
    assert np.array_equal(back['real'], new['real'])
    assert np.array_equal(back['idcpu'], new['idcpu'])
    sim.close()


def test_beam_seams_on_caller_memory():
    """hpb_advance_beam_particles + hpb_beam_shift_slipped with host-side counts (d_np = NULL):
    field-free drift in an external focusing field vs the oracle's advance_beam_slice."""
    import torch
    import hipace_b200 as hp
    from oracle import hipace_oracle as O
    nx = ny = 32
    geom = O.Geometry(nx, ny, 8, (-4., -4., -1.), (4., 4., 1.))
    pc = O.PhysConst.make(True)
    g = hp.NGUARD
    rng = np.random.default_rng(11)
    n = 1000
    islice = 3
    min_z = geom.lo[2] + islice * geom.dz
    bs = dict(x=rng.uniform(-3, 3, n), y=rng.uniform(-3, 3, n),
              z=min_z + rng.uniform(0.0, geom.dz, n), w=rng.uniform(0.5, 1.5, n),
              ux=rng.normal(0, 0.5, n), uy=rng.normal(0, 0.5, n), uz=rng.uniform(1.5, 40., n),
              id=np.arange(1, n + 1, dtype=np.int64), valid=rng.random(n) > 0.1,
              nsub=np.zeros(n, dtype=np.int64), np=n)
    exprs = ['0.3*x', '0.3*y+0.01*t', '0.05*z', '0.', '0.02*x', '0.']
    env = {}
    beam = O.Beam(name='b', charge=-1., mass=1., ppc=(1, 1, 1), profile='flattop', density=1.,
                  zmin=-1, zmax=1, radius=9., n_subcycles=5,
                  external_fields=lambda x, y, z, t: (0.3 * x, 0.3 * y + 0.01 * t, 0.05 * z,
                                                       0. * x, 0.02 * x, 0. * x))
    names = ('Psi', 'Ez', 'Bx', 'By', 'Bz')
    F = {('This', nm): 0.01 * rng.standard_normal((ny + 2 * g, nx + 2 * g)) for nm in names}
    ref = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in bs.items()}
    dt, time = 2.5, 7.0
    O.advance_beam_slice(ref, beam, F, geom, pc, islice, dt, time, 'Absorbing', (-3.5, -3.5), (3.5, 3.5))

    ctx = hp.Context(nx, ny, geom.dx, geom.dy, geom.dz, geom.pos_offset(0), geom.pos_offset(1))
    sl_t = torch.from_numpy(np.stack([F[('This', nm)] for nm in names])).cuda()
    comps = ctx.comps(psi=0, ez=1, bx=2, by=3, bz=4)
    reals = torch.from_numpy(np.stack([bs[k] for k in ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz')])).cuda()
    from hipace_b200.pipeline import pack_idcpu, unpack_idcpu
    idcpu = torch.from_numpy(pack_idcpu(bs['id'], bs['valid']).view(np.int64)).cuda()
    nsub = torch.zeros(n, dtype=torch.int32, device='cuda')
    nblk = (n + 255) // 256
    cls = torch.zeros(2 * nblk, dtype=torch.int32, device='cuda')
    csum = torch.zeros(9, dtype=torch.float64, device='cuda')
    bm = ctx.beam_view(reals, idcpu)
    ext = ctx.extfields(exprs)
    ctx.advance_beam_particles(bm, nsub, ctx.slice_view(sl_t), -1., 1., comps, n_subcycles=5, dt=dt,
                               time=time, min_z=min_z, bc='Absorbing', bc_lo=(-3.5, -3.5),
                               bc_hi=(3.5, 3.5), ext=ext, class_counts=cls, checksum=csum)
    torch.cuda.synchronize()
    out = reals.cpu().numpy()
    ids, valid = unpack_idcpu(idcpu.cpu().numpy().view(np.uint64))
    assert np.array_equal(valid, ref['valid'])
    assert np.array_equal(nsub.cpu().numpy()[valid], ref['nsub'][valid])
    for k, nm in enumerate(('x', 'y', 'z', 'w', 'ux', 'uy', 'uz')):
        err = np.abs(out[k][valid] - ref[nm][valid]).max() / np.abs(ref[nm][valid]).max()
        assert err <= 1e-12, (nm, err)
    assert abs(csum[8].item() - n) == 0 and abs(csum[0].item() - np.abs(bs['x']).sum()) < 1e-9
    # re-binning: stable, invalid dropped, slipped appended behind the next slice's particles
    stay_r = torch.zeros((7, n), dtype=torch.float64, device='cuda')
    stay_i = torch.zeros(n, dtype=torch.int64, device='cuda')
    stay_c = torch.zeros(2, dtype=torch.int64, device='cuda')
    n_next0 = 17
    next_r = torch.zeros((7, n + n_next0), dtype=torch.float64, device='cuda')
    next_i = torch.zeros(n + n_next0, dtype=torch.int64, device='cuda')
    next_c = torch.tensor([n_next0, n_next0], dtype=torch.int64, device='cuda')
    next_ns = torch.full((n + n_next0,), -1, dtype=torch.int32, device='cuda')
    ovf = torch.zeros(1, dtype=torch.int32, device='cuda')
    ctx.beam_shift_slipped(bm, nsub, min_z, cls, ctx.beam_view(stay_r, stay_i), stay_c,
                           ctx.beam_view(next_r, next_i, next_c), next_c, next_ns, ovf)
    torch.cuda.synchronize()
    stay = ref['valid'] & (ref['z'] >= min_z)
    slip = ref['valid'] & ~stay
    assert slip.sum() > 0 and stay.sum() > 0
    assert stay_c.tolist() == [int(stay.sum())] * 2
    assert next_c.tolist() == [n_next0, n_next0 + int(slip.sum())]
    assert ovf.item() == 0
    sid, _ = unpack_idcpu(stay_i.cpu().numpy().view(np.uint64)[:int(stay.sum())])
    assert np.array_equal(sid, ref['id'][stay])
    nid, _ = unpack_idcpu(next_i.cpu().numpy().view(np.uint64)[n_next0:n_next0 + int(slip.sum())])
    assert np.array_equal(nid, ref['id'][slip])
    assert np.array_equal(next_ns.cpu().numpy()[n_next0:n_next0 + int(slip.sum())], ref['nsub'][slip])
    assert np.array_equal(stay_r.cpu().numpy()[2][:int(stay.sum())], out[2][stay])
    ctx.close()
