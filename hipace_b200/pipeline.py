"""Host-side logic of the time-step pipeline (the role of MultiBuffer, src/utils/MultiBuffer.cpp).

Backend-agnostic pieces that the CUDA path (csrc/pipeline.cu: NCCL point-to-point per slice) and
the world_size-2 gloo tests on CPU share:

  * which rank owns which time step (Hipace.cpp:401: `for step = rank; step <= max_step; step += R`)
  * who is upstream / downstream on the ring, and which steps are handed on (Hipace.cpp:441-443)
  * the wire layout of one beam-slice message (MultiBuffer.cpp:611-728: metadata, idcpu, then the
    real components -- here with a fixed per-slice capacity so that counts never visit the host)
  * the exchange of the per-edge ncclUniqueId bytes through torch.distributed

`pack_slice` / `unpack_slice` are the NumPy statement of what csrc/beam.cu's partition kernel
writes and what the receiving rank's kernels read.
"""
from __future__ import annotations

import numpy as np

HEADER_BYTES = 64
REAL_COMPS = ('x', 'y', 'z', 'w', 'ux', 'uy', 'uz')      # BeamIdx order, BeamParticleContainer.H
ID_VALID_BIT = np.uint64(1) << np.uint64(63)


def owned_steps(rank: int, world: int, max_step: int):
    return range(rank, max_step + 1, world)


def upstream(rank: int, world: int) -> int:
    return (rank - 1 + world) % world


def downstream(rank: int, world: int) -> int:
    return (rank + 1) % world


def receives(step: int, world: int) -> bool:
    """step 0 starts from the initial beam; every later step gets it from the step before"""
    return world > 1 and step > 0


def sends(step: int, world: int, max_step: int) -> bool:
    return world > 1 and step + 1 <= max_step


def edge_order(rank: int, world: int):
    """Order in which a rank touches its two ring edges whenever the operation BLOCKS THE HOST until
    the peer of the edge enters the same call (ncclCommInitRank, the first ncclSend / ncclRecv of a
    communicator, ncclCommDestroy): increasing edge index, edge e = (rank e -> rank e+1 mod world).
    Returns [('send' | 'recv', edge), ...]; csrc/pipeline.cu follows exactly this rule.  With any
    other order two ranks can each sit in a blocking call the other one has not reached yet (the
    first 2-GPU run of this code dead-locked that way)."""
    e_send, e_recv = rank, (rank - 1 + world) % world
    ops = [('send', e_send), ('recv', e_recv)]
    return sorted(ops, key=lambda o: o[1])


def message_bytes(capacity: int) -> int:
    return HEADER_BYTES + 64 * capacity


def pack_idcpu(ids, valid):
    """AMReX >= 24 packing: validity in the top bit, id << 24, cpu (= MR level) 0"""
    out = np.asarray(ids, dtype=np.uint64) << np.uint64(24)
    return np.where(np.asarray(valid, dtype=bool), out | ID_VALID_BIT, out)


def unpack_idcpu(idcpu):
    idcpu = np.asarray(idcpu, dtype=np.uint64)
    return ((idcpu & ~ID_VALID_BIT) >> np.uint64(24)).astype(np.int64), (idcpu >> np.uint64(63)) != 0


def pack_slice(bs: dict, capacity: int) -> np.ndarray:
    """bs: x y z w ux uy uz (float64[n]), id (int64[n]), valid (bool[n]) -> uint8[message_bytes]"""
    n = int(bs['x'].size)
    if n > capacity:
        raise ValueError(f'slice of {n} particles exceeds the packet capacity {capacity}')
    buf = np.zeros(message_bytes(capacity), dtype=np.uint8)
    hdr = buf[:HEADER_BYTES].view(np.int64)
    hdr[0] = n
    hdr[1] = n
    body = buf[HEADER_BYTES:]
    body[:8 * capacity].view(np.uint64)[:n] = pack_idcpu(bs['id'], bs['valid'])
    for k, nm in enumerate(REAL_COMPS):
        body[8 * capacity * (k + 1):8 * capacity * (k + 2)].view(np.float64)[:n] = bs[nm]
    return buf


def unpack_slice(buf: np.ndarray, capacity: int) -> dict:
    assert buf.size == message_bytes(capacity)
    n = int(buf[:HEADER_BYTES].view(np.int64)[0])
    body = buf[HEADER_BYTES:]
    ids, valid = unpack_idcpu(body[:8 * capacity].view(np.uint64)[:n].copy())
    bs = {nm: body[8 * capacity * (k + 1):8 * capacity * (k + 2)].view(np.float64)[:n].copy()
          for k, nm in enumerate(REAL_COMPS)}
    bs['id'], bs['valid'] = ids, valid
    bs['nsub'] = np.zeros(n, dtype=np.int64)       # never communicated: restarts at 0
    bs['np'] = n
    return bs


def exchange_edge_ids(dist, rank: int, world: int, my_id: bytes) -> list:
    """all ranks learn the id of every edge e -> e+1 (created by rank e); any backend"""
    import torch
    n = len(my_id)
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    mine = torch.tensor(list(my_id), dtype=torch.uint8, device=dev)
    allv = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allv, mine)
    return [bytes(t.cpu().tolist()) for t in allv]


class HostPipeline:
    """The slice hand-off protocol with host buffers over torch.distributed send/recv (any
    backend): the statement of what csrc/pipeline.cu does with NCCL and device packets, used to
    test the N > 1 path on CPU (gloo) with any engine that exposes

        begin_step(step), solve_one_slice(islice), nz,
        put_beam_slice(ibeam, islice, bs)     install a received slice (before it is first used)
        take_beam_slice(ibeam, islice) -> bs  the pushed slice after solve_one_slice(islice)
        set_time(t) / next_time()             optional: the physical time of a step arrives from the rank
                                              that owns the step before it (MultiBuffer::get_time /
                                              put_time, MultiBuffer.cpp:611-651; Hipace.cpp:411, :445-447)

    Order per slice (Hipace.cpp:582-585, 639-642, 716): slice nz-1 is received before the slice
    loop, slice islice-1 before slice islice is solved (it is the Next slice of the Bx/By
    source), and slice islice is sent right after it was pushed and re-binned.
    """

    def __init__(self, dist, rank: int, world: int, nbeams: int, capacity: int):
        self.dist, self.rank, self.world = dist, rank, world
        self.nbeams, self.capacity = nbeams, capacity
        self.up, self.down = upstream(rank, world), downstream(rank, world)
        self._pending = []

    def _recv(self, engine, islice):
        import torch
        for ib in range(self.nbeams):
            buf = torch.empty(message_bytes(self.capacity), dtype=torch.uint8)
            self.dist.recv(buf, src=self.up)
            engine.put_beam_slice(ib, islice, unpack_slice(buf.numpy(), self.capacity))

    def _send(self, engine, islice):
        import torch
        for ib in range(self.nbeams):
            buf = torch.from_numpy(pack_slice(engine.take_beam_slice(ib, islice), self.capacity))
            # non-blocking: the downstream rank may be a whole step behind
            self._pending.append((self.dist.isend(buf, dst=self.down), buf))

    def run(self, engine, max_step: int):
        nz = engine.nz
        for step in owned_steps(self.rank, self.world, max_step):
            rx, tx = receives(step, self.world), sends(step, self.world, max_step)
            if rx and hasattr(engine, 'set_time'):
                import torch
                t = torch.zeros(1, dtype=torch.float64)
                self.dist.recv(t, src=self.up)
                engine.set_time(float(t[0]))
            engine.begin_step(step)
            if tx and hasattr(engine, 'next_time'):
                import torch
                t = torch.tensor([engine.next_time()], dtype=torch.float64)
                self._pending.append((self.dist.isend(t, dst=self.down), t))
            if rx:
                self._recv(engine, nz - 1)
            for isl in range(nz - 1, -1, -1):
                if rx and isl > 0:
                    self._recv(engine, isl - 1)
                engine.solve_one_slice(isl)
                if tx:
                    self._send(engine, isl)
            if hasattr(engine, 'end_step'):
                engine.end_step(step)
        for req, _ in self._pending:
            req.wait()
        self._pending = []
