#!/usr/bin/env python3
"""bench.py -- slices/s of the per-zeta-slice quasi-static PIC loop on the blowout_wake deck.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (N ranks: time-step pipeline)
  python bench.py --impl reference --steps K --warmup W    CPU restatement of the reference path

A "step" is one whole time step = nz zeta-slices of the deck (Hipace::Evolve body,
src/Hipace.cpp:401-507): plasma re-initialisation + neutralising background + nz x SolveOneSlice.
Workload = BASELINE.json configs[2] (the configuration the metric is quoted on):
blowout_wake, normalised units, 1024 x 1024 transverse x 1024 zeta, ppc 4 (plasma.ppc = 2 2),
fixed_ppc gaussian beam, explicit Bx/By solver, order-2 shapes.  Synthetic, deterministic, no RNG.

Timed region of `value`: K time steps, inputs resident in HBM, CUDA events on the simulation
stream (hpb_sim_evolve records them), barrier + synchronize on both sides, max over ranks.
`e2e`: the same K steps through the public API with HOST buffers -- every step uploads the
whole beam from pinned host memory (the role of MultiBuffer::get_data) and reads back the beam
and the field checksums -- timed by wall clock around the calls.
Working set per slice (177 MB slice array + 419 MB particles) exceeds the 126 MB L2, so no
explicit L2 flush is needed between timed iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The pipeline hands one small packet per slice to the next GPU.  NCCL's default gives a point-to-point
# kernel one CTA per channel on many channels; those CTAs sit on SMs spinning until the peer arrives and
# take issue slots from the slice kernels.  One channel carries the ~100 KB packets with room to spare
# (must be set before NCCL reads its parameters, i.e. before torch.distributed creates a communicator).
os.environ.setdefault('NCCL_MAX_P2P_NCHANNELS', '1')
os.environ.setdefault('NCCL_MIN_P2P_NCHANNELS', '1')

METRIC = 'slices_per_sec'
UNIT = 'slices/s'


def deck_and_overrides(nxy, nz, ppc):
    if (nxy, nz, ppc) == WORKLOADS['configs3'][:3]:
        # laser_blowout_wake_explicit in SI units (the reference's golden deck of that name, scaled up):
        # no beam, a laser pulse drives the wake, the envelope is advanced every slice (multigrid solver,
        # the reference's default) and its slices travel through the pipeline
        meta = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'laser_blowout_wake_explicit.SI.1Rank.json')))
        deck = open(os.path.join(ROOT, meta['deck'])).read()
        ov = dict(meta['overrides'])
        ov.pop('max_step', None)
        ov.update({'amr.n_cell': f'{nxy} {nxy} {nz}', 'plasma.ppc': f'{ppc} {ppc}'})
        return deck, ov
    if (nxy, nz, ppc) == WORKLOADS['configs4'][:3]:
        # the shape of the reference's production deck: two mobile species (ion motion), ppc 9 each
        deck = open(os.path.join(ROOT, 'examples', 'ion_motion_normalized.in')).read()
        # (fixed_ppc with one particle per cell in z puts a whole beam slice at ONE z: all of it slips into
        # the next slice in the same time step, so a slice packet transiently holds two slices' worth --
        # the packet capacity is set accordingly, and the beam radius halved to keep the two rings at
        # 2 x 2048 x 80 000 x 64 B = 21 GB; the plasma, which is what this configuration is about, is unchanged)
        return deck, {'amr.n_cell': f'{nxy} {nxy} {nz}', 'elec.ppc': f'{ppc} {ppc}', 'ions.ppc': f'{ppc} {ppc}',
                      'beam.radius': 0.8, 'beam.slice_capacity': 80000}
    deck = open(os.path.join(ROOT, 'examples', 'blowout_wake_normalized.in')).read()
    ov = {'amr.n_cell': f'{nxy} {nxy} {nz}', 'plasma.ppc': f'{ppc} {ppc}'}
    return deck, ov


# the benchmark line is BASELINE configs[2]; the others are extra lines (profiles/) and parity cases
WORKLOADS = {'configs2': (1024, 1024, 2, 'BASELINE configs[2]'),
             'configs1': (256, 512, 2, 'BASELINE configs[1]'),
             'n1023': (1023, 1024, 2, 'configs[2] on the reference\'s recommended 2^n - 1 grid'),
             'configs3': (512, 1024, 1, 'BASELINE configs[3]: laser_blowout_wake_explicit SI with the envelope solve'),
             'configs4': (2048, 2048, 3, 'BASELINE configs[4] shape: ion motion, two species, ppc 9, normalised units')}


def workload_name(nxy, nz, ppc):
    tag = next((t for n, z, p, t in WORKLOADS.values() if (n, z, p) == (nxy, nz, ppc)), 'custom size')
    deck = {WORKLOADS['configs3'][:3]: 'laser_blowout_wake_explicit SI',
            WORKLOADS['configs4'][:3]: 'ion_motion (two mobile species) normalized'}.get((nxy, nz, ppc),
                                                                                       'blowout_wake_explicit normalized')
    return f'{deck} {nxy}x{nxy}x{nz} ppc={ppc * ppc} ({tag})'


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu):
        self.gpu = gpu
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                       '-lms', '200', '-i', str(self.gpu)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(',') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, smax = [], set(), None
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in rows:
            try:
                r = [t.strip() for t in r]
                sm.append(float(r[1]))
                smax = float(r[2])
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        if sm:
            sm.sort()
            out = {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                   'samples': len(sm)}
        return out


# ------------------------------------------------------------------------------------------------
# CPU restatement (oracle) -- cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_arm(nxy, nz, ppc, steps, warmup, sample_slices):
    """Time the CPU restatement of the reference algorithm (oracle/) on this box's host cores.
    Each step is a bounded sample: `sample_slices` consecutive slices of the same deck starting
    at the beam head (the reference cannot be built here: AMReX/FFTW/MPI absent)."""
    from oracle import cport
    # all physical host cores, whatever OMP_NUM_THREADS the launcher exported (torchrun sets 1)
    cport.set_threads(cport.physical_cores())
    deck, ov = deck_and_overrides(nxy, nz, ppc)
    sim = cport.Simulation(deck, ov)
    sim.begin_step()
    isl = sim.nz - 1 - sim.first_beam_slot()       # start where the beam begins: representative work
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for _ in range(sample_slices):
            sim.solve_one_slice(isl)
            isl -= 1
        times.append(time.perf_counter() - t0)
    tt = sum(times[warmup:])
    return dict(value=steps * sample_slices / tt, seconds=tt, cores=sim.threads,
                sample=f'{sample_slices} slices/step x {steps} steps of the {nxy}x{nxy} ppc={ppc * ppc} '
                       f'deck from the beam head (slice {sim.nz - 1 - sim.first_beam_slot()} down), '
                       f'C/OpenMP port of the oracle, {sim.threads} threads'), tt / steps


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference', 'cufft_ref', 'naive'],
                    help='native: the product path.  reference: the CPU restatement on the host cores.  '
                         'cufft_ref: the product path with the Poisson stage replaced by the reference\'s '
                         'FFTPoissonSolverDirichletFast sequence on cuFFT (csrc/ref_gpu_arm.cu).  naive: a '
                         'straight restatement of the reference\'s GPU algorithm on this box -- one thread '
                         'per particle with plain fp64 atomics and per-particle gathers (generic-order '
                         'kernels), cuFFT Poisson, the reference call order (no fusion, no side stream, no '
                         'programmatic dependent launch)')
    ap.add_argument('--nxy', type=int, default=1024)
    ap.add_argument('--nz', type=int, default=1024)
    ap.add_argument('--ppc', type=int, default=2, help='per direction (2 -> ppc 4)')
    ap.add_argument('--dt', type=float, default=None,
                    help='hipace.dt in 1/omega_p (> 0: the beam evolves from step to step, particles slip '
                         'between slices and every pipeline packet differs from the last)')
    ap.add_argument('--workload', default=None, choices=sorted(WORKLOADS),
                    help='named size (overrides --nxy/--nz/--ppc); the default line is configs2')
    ap.add_argument('--opt', action='append', default=[], metavar='KEY=VALUE',
                    help='hpb_sim_set_option switch for A/B runs (e.g. order=9, fuse=0); recorded in config')
    ap.add_argument('--reorder-period', type=int, default=0,
                    help='plasmas.reorder_period: sort the plasma by cell every N slices (0: never, the default)')
    ap.add_argument('--no-verify', action='store_true', help='skip the N > 1 pipeline-vs-single-GPU check')
    ap.add_argument('--cpu-sample-slices', type=int, default=4)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-profile', action='store_true')
    args = ap.parse_args()

    if args.workload:
        args.nxy, args.nz, args.ppc = WORKLOADS[args.workload][:3]
    if args.dt is None:
        # normalised decks: 1 / omega_p; the SI laser deck: 4 / omega_p of its 2e24 m^-3 plasma in seconds
        # (configs4: the oblique gamma ~ 100 beam of the ion-motion deck slips 0.024 c/omega_p per unit of
        # dt and dz is 0.0059: with dt = 0.05 a fifth of a slice's particles slips per step, which the
        # fixed-capacity slice packets -- initial maximum + 25 % -- hold; larger steps need
        # <beam>.slice_capacity raised, at 64 B x nz x capacity per ring)
        key = (args.nxy, args.nz, args.ppc)
        args.dt = {WORKLOADS['configs3'][:3]: 4.0 * 10.e-6 / 299792458., WORKLOADS['configs4'][:3]: 0.05}.get(key, 1.0)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    # a stuck rank must say where it is stuck and leave, never hold the box until an outer limit
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('HPB_BENCH_WATCHDOG', '1500')), exit=True)
    log = (lambda m: print(f'[bench rank {rank}] {m}', file=sys.stderr, flush=True)) if os.environ.get('HPB_BENCH_LOG') \
        else (lambda m: None)
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    K, W = args.steps, max(args.warmup, 0)
    cfg = {'workload': workload_name(args.nxy, args.nz, args.ppc), 'nx': args.nxy, 'ny': args.nxy,
           'nz': args.nz, 'ppc': args.ppc * args.ppc,
           'units': 'SI' if (args.nxy, args.nz, args.ppc) == WORKLOADS['configs3'][:3] else 'normalized', 'dt': args.dt,
           'solver': 'explicit (FFT/DST Poisson x3 + multigrid BxBy)',
           'l2_policy': 'working set per slice (~600 MB) > L2 (126 MB); no flush needed',
           'parallelism': f'time-step pipeline x{world}' if world > 1 else 'single GPU'}

    if args.impl == 'reference':
        if rank != 0:
            return 0
        base, sec_per_step = cpu_arm(args.nxy, args.nz, args.ppc, K, max(W, 1), args.cpu_sample_slices)
        line = {'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT,
                'n_gpus': args.gpus, 'steps': K, 'warmup': W, 'ms_per_step': 1e3 * sec_per_step,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
                'data': 'synthetic', 'config': cfg,
                'cpu_baseline': {'value': base['value'], 'unit': UNIT, 'cores': base['cores'],
                                 'kind': 'port', 'sample': base['sample']},
                'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import hipace_b200 as hp

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    deck, ov = deck_and_overrides(args.nxy, args.nz, args.ppc)
    if args.dt:
        ov['hipace.dt'] = args.dt
    if args.reorder_period:
        ov['plasmas.reorder_period'] = args.reorder_period
        cfg['reorder_period'] = args.reorder_period
    for kv in list(args.opt):          # process-wide switches take effect when a context is created
        if kv.split('=', 1)[0] in ('bluestein_min_prime',):
            hp.set_global_option(kv.split('=', 1)[0], float(kv.split('=', 1)[1]))
    n_beams = int(hp.deck_check(deck, ov)['n_beams'])
    if n_beams == 0:
        args.no_e2e = True          # the end-to-end leg moves the beam through host memory: nothing to move
    sim = hp.Simulation(deck, ov, device=local_rank)
    sim.set_option('checksums', 0)
    arm_opts = {'cufft_ref': ['poisson_impl=1'],
                'naive': ['poisson_impl=1', 'generic_order_kernels=1', 'fuse=0', 'side_stream=0', 'pdl=0']}
    args.opt = arm_opts.get(args.impl, []) + args.opt
    for kv in args.opt:
        k, v = kv.split('=', 1)
        sim.set_option(k, float(v))
    if args.opt:
        cfg['options'] = list(args.opt)
    sim.pipeline_init(rank, world, dist)

    # ---- device-resident leg ---------------------------------------------------------------
    # rank r owns the time steps r, r + world, ... (Hipace.cpp:401); every rank runs `count`
    # steps (weak scaling).  A run is self-contained: its last step hands nothing on, so the
    # pipeline is drained when run() returns and the barriers around the timed region are safe.
    def run_steps(count):
        st = sim.run(count * world - 1, rank, world)
        return st['slice_loop_ms'], st['n_kernel_launches'], st

    log('warm-up')
    run_steps(W)
    barrier()
    log('timed region')
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    sim.timer_start()
    loop_ms, launches, st_last = run_steps(K)
    dev_ms = sim.timer_stop()          # CUDA events on the simulation stream around the K steps
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device='cuda')
    tl = torch.tensor([float(launches)], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    launches = int(tl[0])
    job_ms = dev_ms_max               # whole job = the slowest rank's device time (incl. plasma init)
    total_slices = K * args.nz * world
    value = total_slices / (job_ms * 1e-3)
    n_pushed = st_last['n_plasma_pushed'] + st_last['n_beam_pushed']
    ns_per_push = dev_ms * 1e6 / max(n_pushed, 1) / world

    # ---- end-to-end leg (host buffers) -----------------------------------------------------
    # N = 1: every step uploads its input beam from pinned host memory (MultiBuffer::get_data
    # with host staging) and reads back the pushed beam + the field checksums.
    # N > 1: the beam of step 0 comes from pinned host memory; later steps receive theirs from
    # the upstream GPU over NVLink (that IS the public multi-GPU API); every step still reads
    # back its pushed beam + checksums (the per-step diagnostic of a production run).
    log('end-to-end leg')
    e2e = None
    if not args.no_e2e:
        npb = sim.beam_np()
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
        host = {'real': pin((7, max(npb, 1)), torch.float64), 'idcpu': pin((max(npb, 1),), torch.int64).view(np.uint64),
                'slot_off': np.zeros(args.nz + 1, dtype=np.int64)}
        sim.get_beam(host)
        sim.set_option('checksums', 1)
        beam_bytes = 7 * 8 * npb + 8 * npb
        ncs = len(sim.checksums())
        from hipace_b200 import pipeline as pl

        def e2e_run(count):
            max_step = count * world - 1
            sim.set_option('max_step', max_step)
            for step in pl.owned_steps(rank, world, max_step):
                if world == 1 or step == 0:
                    sim.set_beam(host)             # H2D: this step's beam from pinned host memory
                sim.evolve(step, step)             # plasma init + nz slices; D2H of the checksums
                sim.get_beam(host)                 # D2H: the beam after the step
        e2e_run(1)
        barrier()
        t0 = time.perf_counter()
        e2e_run(K)
        barrier()
        e2e_s = time.perf_counter() - t0
        nsteps = K * world
        h2d = beam_bytes if world == 1 else beam_bytes / nsteps
        e2e = {'value': nsteps * args.nz / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
               'd2h_bytes_per_step': beam_bytes + 8 * ncs, 'ms_per_step': 1e3 * e2e_s / K,
               'api': 'hipace_b200.Simulation.set_beam/evolve/get_beam -> hpb_sim_* C-ABI'
                      + ('' if world == 1 else '; steps > 0 receive their beam from the upstream GPU (NCCL p2p)')}
        sim.set_option('checksums', 0)

    # ---- N > 1: the pipeline's result against a single-GPU run of the same steps (untimed) -----
    # Steps 0 .. world-1 through the ring (one per rank, dt > 0: the beam rank r receives is the one
    # rank r-1 pushed), then the rank that owns the last step replays all `world` steps alone on its
    # own GPU: beam moments and field checksums of the last step must agree to 1e-9.
    log('verification leg')
    verify = None
    if world > 1 and not args.no_verify:
        sim.set_option('checksums', 1)
        if n_beams:
            sim.set_option('beam_from_host', 0)     # the e2e leg uploaded a beam: step 0 starts from the deck's again
        sim.run(world - 1, rank, world)
        barrier()
        if rank == world - 1:
            got_f, got_b = sim.checksums(), (sim.beam_checksums(0) if n_beams else {})
            solo = hp.Simulation(deck, ov, device=local_rank)
            solo.pipeline_init(0, 1, None)
            solo.set_option('max_step', world - 1)
            want_f = solo.evolve(0, world - 1)
            want_b = solo.beam_checksums(0) if n_beams else {}
            solo.close()
            worst = 0.
            for want, got in ((want_f, got_f), (want_b, got_b)):
                for k, w in want.items():
                    err = abs(got[k] - w) / max(abs(w), 1e-300) if w != 0. else abs(got[k])
                    worst = max(worst, err)
            verify = {'steps': world, 'rank': rank, 'worst_rel_err': worst, 'ok': bool(worst <= 1e-9),
                      'checked': sorted(want_f) + ['beam:' + k for k in sorted(want_b)]}
        sim.set_option('checksums', 0)
        barrier()
        if dist is not None:
            box = [verify]
            dist.broadcast_object_list(box, src=world - 1)
            verify = box[0]

    # ---- per-stage profile + roofline of the dominant kernel (extra, untimed pass) -----------
    log('profile leg')
    roofline = None
    stages = None
    if not args.no_profile and rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
        sim.set_option('profile', 1)
        # a lone step: nothing is handed downstream (the laser deck keeps max_step = 1 on a single GPU so
        # that the envelope solve is part of the profiled slice)
        sim.set_option('max_step', 1 if (n_beams == 0 and world == 1) else 0)
        sim.evolve(0, 0)
        sp = sim.stats()
        sim.set_option('profile', 0)
        nsl = sp['n_slices']
        P = sp['n_plasma_pushed'] / nsl
        Gc = (args.nxy + 4) ** 2
        # algorithmic bytes per launch (particle kernels) / per slice (solver stages), SURVEY.md 8(d)
        Nc = args.nxy ** 2
        ncyc = sp['n_mg_vcycles'] / nsl
        alg = {'deposit': 56 * P + 4 * 8 * Gc, 'explicit': 56 * P + 8 * 8 * Gc, 'push': 128 * P + 5 * 8 * Gc,
               'poisson': 3 * 6 * 8 * Nc,                       # 3 solves x practical 3-pass model
               'mg': (30 * ncyc + 17) * 8 * Gc}                 # hpmg finest-level model
        stages = {k: sp['ms_' + k] / nsl for k in ('deposit', 'poisson', 'explicit', 'mg', 'push', 'other')}
        stages['mg_vcycles_per_slice'] = ncyc
        kernel_of = {'deposit': 'k_deposit_current', 'explicit': 'k_explicit_deposition',
                     'push': 'k_advance_plasma'}
        # default driver: ::DepositCurrent of the next slice is fused into the push kernel (the
        # 56 B/particle re-read disappears, w is read on top of the push streams)
        fused_deposit = stages['deposit'] < 0.02 * stages['push']
        if fused_deposit:
            alg['push'] += 8 * P + 4 * 8 * Gc
            del alg['deposit'], kernel_of['deposit']
        dom = max(kernel_of, key=lambda k: stages[k])
        ach = alg[dom] / (stages[dom] * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
            if (tr.get('nxy'), tr.get('ppc')) == (args.nxy, args.ppc * args.ppc):
                traffic = tr['kernels'].get(kernel_of[dom])
                traffic_src = tr.get('source')
        except Exception:
            pass
        # second roof: fp64.  flops per particle = 2 DFMA + DMUL + DADD thread instructions of the ncu
        # op counters (smsp__sass_thread_inst_executed_op_d{fma,mul,add}_pred_on, profiles/r02/r02F_particles_raw.csv:
        # the round-2 default kernels, slice 623) divided by the particles of the launch; the peak is measured on
        # this box (csrc/peaks.cu)
        flops_pp = {'push': 1021.2, 'explicit': 496.3}
        try:
            fp64_peak = hp.measure_fp64_peak(local_rank)
        except Exception:
            fp64_peak = None
        fp64 = {}
        for k, fpp in flops_pp.items():
            if k in stages and stages[k] > 0 and fp64_peak:
                tf = fpp * P / (stages[k] * 1e-3) / 1e12
                fp64[k] = {'achieved': tf, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': tf / fp64_peak,
                           'flops_per_particle': fpp}
        roofline = {'kernel': kernel_of[dom],
                    'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': alg[dom], 'avg_launch_ms': stages[dom],
                    'deposit_fused_into_push': fused_deposit,
                    'fp64': fp64, 'fp64_peak_source': 'measured on this box: DFMA kernel, best of 5 (csrc/peaks.cu)',
                    'all': {k: {'GBps': alg[k] / (stages[k] * 1e-3) / 1e9, 'ms': stages[k],
                                'frac': alg[k] / (stages[k] * 1e-3) / 1e9 / peak} for k in alg}}

    # ---- CPU baseline (rank 0, N = 1) -----------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            base, _ = cpu_arm(args.nxy, args.nz, args.ppc, 2, 1, args.cpu_sample_slices)
            cpu_baseline = {'value': base['value'], 'unit': UNIT, 'cores': base['cores'], 'kind': 'port',
                            'sample': base['sample']}
        except Exception as e:      # the baseline is a report, never a dependency of the GPU path
            cpu_baseline = {'value': None, 'unit': UNIT, 'cores': 0, 'kind': 'port',
                            'sample': f'unavailable: {type(e).__name__}: {e}'}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
                **({'impl': args.impl} if args.impl != 'native' else {}),
                'ms_per_step': job_ms / K, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': cfg,
                'ns_per_particle_step': ns_per_push, 'ms_per_slice': job_ms / (K * args.nz),
                'slice_loop_ms_per_step_rank0': loop_ms / K,
                'gpu_launches': int(launches), 'fused_driver_order': bool(st_last.get('n_fused_slices', 0) > 0),
                'clocks': clocks, 'e2e': e2e, 'roofline': roofline,
                'stage_ms_per_slice': stages, 'cpu_baseline': cpu_baseline, 'pipeline_verify': verify,
                'wall_ms_per_step': wall_ms_max / K}
        print(json.dumps(line))
    barrier()
    sim.close()
    if dist is not None:
        dist.destroy_process_group()
    if verify is not None and not verify.get('ok', False):
        print(f'bench.py: the pipeline result differs from the single-GPU run: {verify}', file=sys.stderr)
        return 1
    return 0


if __name__ == '__main__':
    sys.exit(main())
