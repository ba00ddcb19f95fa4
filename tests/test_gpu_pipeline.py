"""GPU side of the time-step pipeline: the wire message the partition kernel packs, and (when the
box has >= 2 GPUs) two NCCL ranks against a single-GPU run of the same deck."""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OV = {'amr.n_cell': '48 48 20', 'hipace.dt': 4., 'beam.u_mean': '0. 0. 3.', 'beam.ppc': '1 1 2',
      'beam.zmin': -3., 'beam.zmax': 3., 'beam.density': 0.5, 'beam.n_subcycles': 4,
      'beam.radius': 3., 'max_step': 4}


def _deck():
    return open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()


def test_wire_message_matches_host_packing():
    """after 3 steps the packet of every slice == pipeline.pack_slice(oracle slice): header,
    idcpu bits (ids, order, validity) bit-exact, the real components to 1e-9"""
    import hipace_b200 as hp
    from hipace_b200 import pipeline as pl
    from oracle.hipace_oracle import Simulation as Oracle
    ref = Oracle(_deck(), OV)
    ref.evolve(step_end=2)
    sim = hp.Simulation(_deck(), OV)
    sim.evolve(0, 2)
    cap = sim.beam_slice_capacity()
    assert sim.pipeline_message_bytes() == pl.message_bytes(cap)
    nz = sim.nz
    nonempty = 0
    for isl in range(nz):
        got = sim.beam_packet(nz - 1 - isl)
        bs = ref.beams[0].slices[isl]
        n = bs['np']
        want = pl.pack_slice({k: (v[:n] if isinstance(v, np.ndarray) else v) for k, v in bs.items()}, cap)
        assert got[:16].view(np.int64).tolist() == [n, n]
        assert np.array_equal(got[64:64 + 8 * cap].view(np.uint64)[:n], want[64:64 + 8 * cap].view(np.uint64)[:n])
        g, w = pl.unpack_slice(got, cap), pl.unpack_slice(want, cap)
        for k in pl.REAL_COMPS:
            if n:
                assert np.abs(g[k] - w[k]).max() <= 1e-9 * max(np.abs(w[k]).max(), 1e-300), (isl, k)
        nonempty += n > 0
    assert nonempty >= 8
    sim.close()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _laser_case():
    import json
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'laser_blowout_wake_explicit.1Rank.json')))
    deck = open(os.path.join(ROOT, meta['deck'])).read()
    ov = dict(meta['overrides'], **{'amr.n_cell': '32 32 40', 'max_step': 3, 'hipace.dt': 4.})
    return deck, ov


def _laser_worker(rank, world, port, out_dir, solver):
    sys.path.insert(0, ROOT)
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('HPB_TEST_WATCHDOG', '240')), exit=True)
    import torch
    import torch.distributed as dist
    import hipace_b200 as hp
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    deck, ov = _laser_case()
    ov['lasers.solver_type'] = solver
    sim = hp.Simulation(deck, ov, device=rank)
    sim.pipeline_init(rank, world, dist)
    sim.run(ov['max_step'], rank, world)
    pickle.dump({'cs': sim.checksums()}, open(os.path.join(out_dir, f'rank{rank}.pkl'), 'wb'))
    torch.cuda.synchronize()
    dist.barrier()
    sim.close()
    dist.destroy_process_group()


@pytest.mark.parametrize('solver', ['fft', 'multigrid'])
def test_two_gpu_laser_pipeline_matches_single_gpu(solver):
    """the laser envelope slices A^{n+1}, A^n travel with the per-slice hand-over (MultiBuffer.cpp:444-490):
    4 time steps of the laser-driven blow-out deck on 2 ranks == the same steps on one GPU"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    import torch.multiprocessing as mp
    import hipace_b200 as hp
    world = 2
    deck, ov = _laser_case()
    ov['lasers.solver_type'] = solver
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_laser_worker, args=(world, _free_port(), d, solver), nprocs=world, join=True)
        res = [pickle.load(open(os.path.join(d, f'rank{r}.pkl'), 'rb')) for r in range(world)]
    last = res[ov['max_step'] % world]
    sim = hp.Simulation(deck, ov)
    cs = sim.evolve(0, ov['max_step'])
    big = max(abs(v) for v in cs.values())
    for k, v in cs.items():
        assert abs(last['cs'][k] - v) <= 1e-9 * abs(v) + 1e-12 * big, k
    sim.close()


ADAPTIVE_STEPS = 7


def _adaptive_case():
    import json
    meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'adaptive_time_step.1Rank.json')))
    return open(os.path.join(ROOT, meta['deck'])).read(), dict(meta['overrides'], max_step=ADAPTIVE_STEPS)


def _adaptive_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('HPB_TEST_WATCHDOG', '240')), exit=True)
    import torch
    import torch.distributed as dist
    import hipace_b200 as hp
    from hipace_b200 import pipeline as pl
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    deck, ov = _adaptive_case()
    sim = hp.Simulation(deck, ov, device=rank)
    sim.pipeline_init(rank, world, dist)
    sim.set_option('max_step', ADAPTIVE_STEPS)
    times = {}
    for step in pl.owned_steps(rank, world, ADAPTIVE_STEPS):
        sim.evolve(step, step)
        times[step] = sim.time()
    pickle.dump({'cs': sim.checksums(), 'bcs': sim.beam_checksums(), 'times': times},
                open(os.path.join(out_dir, f'rank{rank}.pkl'), 'wb'))
    torch.cuda.synchronize()
    dist.barrier()
    sim.close()
    dist.destroy_process_group()


def test_two_gpu_adaptive_time_step_matches_oracle_emulation():
    """hipace.dt = adaptive over 2 NCCL ranks (the time travels with hpb_pipeline_put_time / get_time =
    MultiBuffer::put_time / get_time, every rank keeps its own dt and min_uz_mq): the sequence of
    (time, dt) of all 8 steps and the last step's field / beam checksums equal the oracle's emulation of
    a 2-rank run (Simulation(numprocs = 2), itself reproduced bit for bit by 2 gloo ranks in
    tests/test_pipeline_gloo.py); the sequence differs from the 1-rank one."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    import torch.multiprocessing as mp
    from oracle.hipace_oracle import Simulation as Oracle
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_adaptive_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        res = [pickle.load(open(os.path.join(d, f'rank{r}.pkl'), 'rb')) for r in range(world)]
    deck, ov = _adaptive_case()
    seq = {}
    for R in (1, world):
        ref = Oracle(deck, ov, numprocs=R)
        seq[R] = []
        begin = ref.begin_step

        def spy(step, _ref=ref, _b=begin, _s=seq[R]):
            _b(step)
            _s.append((_ref.time, _ref.dt))
        ref.begin_step = spy
        want = ref.evolve(step_end=ADAPTIVE_STEPS)
    got = {}
    for r in res:
        got.update(r['times'])
    for step in range(ADAPTIVE_STEPS + 1):
        for a, b in zip(got[step], seq[world][step]):
            assert abs(a - b) <= 1e-12 * abs(b), (step, got[step], seq[world][step])
    assert max(abs(a[1] - b[1]) / b[1] for a, b in zip(seq[1], seq[world])) > 1e-4
    last = res[ADAPTIVE_STEPS % world]
    big = max(abs(v) for v in want.values())
    for k, v in want.items():
        assert abs(last['cs'][k] - v) <= 1e-9 * abs(v) + 1e-12 * big, k
    for k, v in ref.beam_checksums()['beam'].items():
        if k in last['bcs']:
            assert abs(last['bcs'][k] - v) <= 1e-9 * abs(v) + 1e-30, k


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    # a hung rank must never hold the GPU box: dump where it is stuck and exit
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('HPB_TEST_WATCHDOG', '240')), exit=True)
    log = lambda msg: print(f'[rank {rank}] {msg}', file=sys.stderr, flush=True)
    import torch
    import torch.distributed as dist
    import hipace_b200 as hp
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    log('process group up')
    sim = hp.Simulation(_deck(), OV, device=rank)
    sim.pipeline_init(rank, world, dist)
    log('pipeline edges connected')
    sim.run(OV['max_step'], rank, world)
    log('steps done')
    n = max(sim.beam_np(), 1)
    host = {'real': np.zeros((7, n)), 'idcpu': np.zeros(n, dtype=np.uint64),
            'slot_off': np.zeros(sim.nz + 1, dtype=np.int64)}
    sim.get_beam(host)
    pickle.dump({'cs': sim.checksums(), 'beam': host, 'bcs': sim.beam_checksums()},
                open(os.path.join(out_dir, f'rank{rank}.pkl'), 'wb'))
    torch.cuda.synchronize()
    dist.barrier()
    sim.close()
    dist.destroy_process_group()


def test_two_gpu_pipeline_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    import torch.multiprocessing as mp
    import hipace_b200 as hp
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        res = [pickle.load(open(os.path.join(d, f'rank{r}.pkl'), 'rb')) for r in range(world)]
    last = res[OV['max_step'] % world]
    sim = hp.Simulation(_deck(), OV)
    cs = sim.evolve(0, OV['max_step'])
    n = max(sim.beam_np(), 1)
    host = {'real': np.zeros((7, n)), 'idcpu': np.zeros(n, dtype=np.uint64),
            'slot_off': np.zeros(sim.nz + 1, dtype=np.int64)}
    sim.get_beam(host)
    for k, v in cs.items():
        assert abs(last['cs'][k] - v) <= 1e-9 * abs(v) + 1e-30, k
    assert np.array_equal(last['beam']['slot_off'], host['slot_off'])
    assert np.array_equal(last['beam']['idcpu'], host['idcpu'])
    for k in range(7):
        a, b = last['beam']['real'][k], host['real'][k]
        assert np.abs(a - b).max() <= 1e-9 * max(np.abs(b).max(), 1e-300), k
    sim.close()
