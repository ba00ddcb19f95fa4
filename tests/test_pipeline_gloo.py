"""The N > 1 path on CPU: two gloo ranks run the time-step pipeline protocol
(hipace_b200/pipeline.py, the host-side statement of csrc/pipeline.cu) with the oracle as the
compute engine, and must reproduce a single-process run of the same deck bit for bit --
including beam particles that slip from one slice into the next between time steps.
"""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OV = {'amr.n_cell': '24 24 12', 'hipace.dt': 4., 'beam.u_mean': '0. 0. 3.', 'beam.ppc': '1 1 2',
      'beam.zmin': -3., 'beam.zmax': 3., 'beam.density': 0.5, 'beam.n_subcycles': 4,
      'beam.radius': 3., 'max_step': 4}
CAP = 512
KEYS = ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz', 'id')


class OracleEngine:
    def __init__(self, deck, ov, numprocs=1):
        from oracle.hipace_oracle import Simulation
        self.sim = Simulation(deck, ov, numprocs=numprocs)
        self.nz = self.sim.geom.nz

    def end_step(self, step):
        self.sim.end_step_adaptive(step)

    def set_time(self, t):                      # MultiBuffer::get_time
        self.sim._next_time = t

    def next_time(self):                        # MultiBuffer::put_time
        s = self.sim
        return s._next_time if s.adaptive_dt else s.time + s.dt

    def begin_step(self, step):
        self.sim.checksums = {}
        self.sim.begin_step(step)

    def solve_one_slice(self, isl):
        self.sim.solve_one_slice(isl)

    def put_beam_slice(self, ib, isl, bs):
        self.sim.beams[ib].slices[isl] = bs

    def take_beam_slice(self, ib, isl):
        bs = self.sim.beams[ib].slices[isl]
        assert bs['x'].size == bs['np']          # slipped particles were moved on
        return bs


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker_adaptive(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import json
    import torch.distributed as dist
    from hipace_b200 import pipeline as pl
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'adaptive_time_step.1Rank.json')))
    ov = dict(meta['overrides'], max_step=ADAPTIVE_STEPS)
    eng = OracleEngine(open(os.path.join(ROOT, meta['deck'])).read(), ov, numprocs=world)
    times = {}
    begin = eng.begin_step

    def spy(step):
        begin(step)
        times[step] = (eng.sim.time, eng.sim.dt)
    eng.begin_step = spy
    pl.HostPipeline(dist, rank, world, 1, 8192).run(eng, ADAPTIVE_STEPS)
    res = {'times': times, 'checksums': eng.sim.checksums,
           'beam': {i: {k: bs[k][:bs['np']] for k in KEYS} for i, bs in eng.sim.beams[0].slices.items()}}
    pickle.dump(res, open(os.path.join(out_dir, f'rank{rank}.pkl'), 'wb'))
    dist.barrier()
    dist.destroy_process_group()


ADAPTIVE_STEPS = 7


@pytest.mark.parametrize('world', [2, 3])
def test_adaptive_time_step_through_the_pipeline(world):
    """hipace.dt = adaptive over `world` ranks: every rank keeps its own dt / min_uz_mq (computed from
    the beam of ITS previous step, used `world` steps later), the physical time travels with put_time /
    get_time.  `world` gloo ranks must reproduce the serial emulation Simulation(numprocs = world) bit
    for bit -- and that run must differ from the one-rank sequence of time steps (the reference's
    adaptive_time_step.1Rank golden run), otherwise the test would not see the per-rank state."""
    import json
    import torch.multiprocessing as mp
    from oracle.hipace_oracle import Simulation
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker_adaptive, args=(world, _free_port(), d), nprocs=world, join=True)
        res = [pickle.load(open(os.path.join(d, f'rank{r}.pkl'), 'rb')) for r in range(world)]
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'adaptive_time_step.1Rank.json')))
    ov = dict(meta['overrides'], max_step=ADAPTIVE_STEPS)
    text = open(os.path.join(ROOT, meta['deck'])).read()
    seq = {}
    for R in (1, world):
        ref = Simulation(text, ov, numprocs=R)
        seq[R] = []
        begin = ref.begin_step

        def spy(step, _ref=ref, _b=begin, _s=seq[R]):
            _b(step)
            _s.append((_ref.time, _ref.dt))
        ref.begin_step = spy
        ref.evolve(step_end=ADAPTIVE_STEPS)
    got = {}
    for r in res:
        got.update(r['times'])
    assert [got[s] for s in range(ADAPTIVE_STEPS + 1)] == seq[world]
    assert seq[world] != seq[1]
    last = res[ADAPTIVE_STEPS % world]
    for k, v in ref.checksums.items():
        assert last['checksums'][k] == v, k
    for isl, bs in ref.beams[0].slices.items():
        for k in KEYS:
            assert np.array_equal(last['beam'][isl][k], bs[k][:bs['np']]), (isl, k)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from hipace_b200 import pipeline as pl
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ids = pl.exchange_edge_ids(dist, rank, world, bytes([rank + 1]) * 128)
    assert ids == [bytes([r + 1]) * 128 for r in range(world)]
    deck = open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()
    eng = OracleEngine(deck, OV)
    pl.HostPipeline(dist, rank, world, 1, CAP).run(eng, OV['max_step'])
    res = {'steps': list(pl.owned_steps(rank, world, OV['max_step'])),
           'checksums': eng.sim.checksums,
           'beam': {i: {k: bs[k][:bs['np']] for k in KEYS} for i, bs in eng.sim.beams[0].slices.items()}}
    pickle.dump(res, open(os.path.join(out_dir, f'rank{rank}.pkl'), 'wb'))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_multi_rank_pipeline_reproduces_serial_run(world):
    """world = 2: both neighbours of a rank are the same peer; world = 3: the smallest ring whose
    upstream and downstream peers differ (the case 4 and 8 GPUs are)"""
    import torch.multiprocessing as mp
    from oracle.hipace_oracle import Simulation
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        res = [pickle.load(open(os.path.join(d, f'rank{r}.pkl'), 'rb')) for r in range(world)]
    if world == 2:
        assert res[0]['steps'] == [0, 2, 4] and res[1]['steps'] == [1, 3]
    else:
        assert [r['steps'] for r in res] == [[0, 3], [1, 4], [2]]
    deck = open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()
    ref = Simulation(deck, OV)
    ref.evolve(step_end=OV['max_step'])
    last = res[OV['max_step'] % world]             # the rank that ran the last step
    for k, v in ref.checksums.items():
        assert last['checksums'][k] == v, k         # same arithmetic on both sides: bit-exact
    nslip = 0
    for isl, bs in ref.beams[0].slices.items():
        n = bs['np']
        got = last['beam'][isl]
        for k in KEYS:
            assert np.array_equal(got[k], bs[k][:n]), (isl, k)
        nslip += n
    assert nslip > 0


def test_schedule_and_wire_layout():
    from hipace_b200 import pipeline as pl
    # every step is owned exactly once, in ring order
    for world in (1, 2, 3, 8):
        owners = {}
        for r in range(world):
            for s in pl.owned_steps(r, world, 20):
                assert s not in owners
                owners[s] = r
        assert sorted(owners) == list(range(21))
        for s in range(1, 21):
            assert owners[s] == pl.downstream(owners[s - 1], world)
            assert pl.upstream(owners[s], world) == owners[s - 1]
        assert not pl.receives(0, world) and not pl.sends(20, world, 20)
    # message = 64-byte header + idcpu[cap] + 7 real arrays [cap] (MultiBuffer.cpp:611-728)
    rng = np.random.default_rng(0)
    n, cap = 37, 64
    bs = {k: rng.standard_normal(n) for k in pl.REAL_COMPS}
    bs['id'] = np.arange(5, 5 + n)
    bs['valid'] = rng.random(n) > 0.3
    buf = pl.pack_slice(bs, cap)
    assert buf.size == pl.message_bytes(cap) == 64 + 64 * cap
    assert buf[:8].view(np.int64)[0] == n
    assert np.array_equal(buf[64 + 8 * cap * 3:64 + 8 * cap * 4].view(np.float64)[:n], bs['z'])
    back = pl.unpack_slice(buf, cap)
    for k in pl.REAL_COMPS + ('id', 'valid'):
        assert np.array_equal(back[k], bs[k]), k
    assert (back['nsub'] == 0).all()
    with pytest.raises(ValueError):
        pl.pack_slice(bs, 16)


def test_library_exports_every_declared_symbol():
    """include/hpb200.h <-> libhpb200.so (no compute calls: there is no GPU here)"""
    import re
    import hipace_b200 as hp
    hdr = open(os.path.join(ROOT, 'include', 'hpb200.h')).read()
    declared = set(re.findall(r'\b(hpb_[a-z0-9_]+)\s*\(', hdr))
    L = hp.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(hp.EXPORTS), declared ^ set(hp.EXPORTS)


@pytest.mark.parametrize('world', [2, 3, 4, 8])
def test_blocking_edge_order_cannot_deadlock(world):
    """Rendezvous model of the blocking NCCL calls: a rank proceeds past an edge only when the
    other end of that edge is at the same edge.  The ordering rule of pipeline.edge_order must
    drain for every ring size; the naive "send first, then receive" order must not (it is the
    dead-lock the first 2-GPU run hit), which shows the model can tell the difference."""
    from hipace_b200 import pipeline as pl

    def drains(order_of):
        todo = {r: list(order_of(r)) for r in range(world)}
        progress = True
        while progress and any(todo.values()):
            progress = False
            for r in range(world):
                if not todo[r]:
                    continue
                kind, e = todo[r][0]
                peer = (e + 1) % world if kind == 'send' else e        # other end of edge e
                if todo[peer] and todo[peer][0][1] == e and todo[peer][0][0] != kind:
                    todo[r].pop(0); todo[peer].pop(0)
                    progress = True
        return not any(todo.values())

    assert drains(lambda r: pl.edge_order(r, world))
    assert not drains(lambda r: [('send', r), ('recv', (r - 1 + world) % world)])
