"""The C/OpenMP CPU restatement (oracle/hipace_cport.c via oracle/cport.py) is pinned twice:
per cell against the NumPy oracle, and against the reference's golden checksums."""
import json
import os

import numpy as np
import pytest

from oracle import cport
from oracle.hipace_oracle import Simulation as Oracle

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.mark.parametrize('case', ['linear_wake.normalized.1Rank', 'blowout_wake_explicit.2Rank',
                                  'gaussian_linear_wake.normalized.1Rank', 'gaussian_linear_wake.SI.1Rank',
                                  'linear_wake.SI.1Rank'])
def test_cport_matches_reference_golden(case, repo_root):
    meta = json.load(open(os.path.join(GOLD, case + '.json')))
    deck = open(os.path.join(repo_root, meta['deck'])).read()
    ov = dict(meta['overrides'])
    ov.pop('max_step', None)
    sim = cport.Simulation(deck, ov)
    cs = sim.evolve()
    for name, want in meta['checksums']['lev=0'].items():
        assert abs(cs[name] - want) <= 1e-9 * abs(want) + 1e-40, (name, cs[name], want)
    assert sim.n_qsa_violation == 0


@pytest.mark.parametrize('ov,nsl', [({}, 45), ({'amr.n_cell': '63 63 100', 'plasma.ppc': '2 2'}, 30)])
def test_cport_matches_numpy_oracle_per_cell(ov, nsl, repo_root):
    deck = open(os.path.join(repo_root, 'examples', 'blowout_wake_normalized.in')).read()
    a, b = cport.Simulation(deck, ov), Oracle(deck, ov)
    a.begin_step()
    b.begin_step()
    nz = b.geom.nz
    for isl in range(nz - 1, nz - 1 - nsl, -1):
        a.solve_one_slice(isl)
        b.solve_one_slice(isl)
        assert a.mg_cycles[-1] == b.mg_cycles[-1] or isl > nz - 4
    for k, want in b.F.items():
        got = a.F[k]
        err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)
        assert err <= 1e-10, (k, err)
    pa, pb = a.plasmas[0], b.plasmas[0]
    assert np.array_equal(pa.valid, pb.valid)
    for nm in ('x', 'y', 'ux', 'uy', 'psi', 'x_prev', 'ux_half', 'psi_half', 'w'):
        va, vb = getattr(pa, nm), getattr(pb, nm)
        assert np.abs(va - vb).max() <= 1e-10 * max(np.abs(vb).max(), 1e-6), nm
