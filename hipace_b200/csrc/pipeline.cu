// Time-step pipeline over the GPUs of one box: the role of MultiBuffer (src/utils/MultiBuffer.cpp)
// with MPI replaced by NCCL point-to-point over NVLink.
//
// Rank r owns the time steps r, r + R, ... (Hipace.cpp:401).  The only state that flows between
// time steps is the beam: for every slice, rank r receives the slice packet that rank r-1 pushed
// in the previous time step and sends its own pushed packet to rank r+1 (ring).  There is no
// collective; besides the beam only the physical time of the next step travels (hpb_pipeline_put_time /
// get_time = MultiBuffer::put_time / get_time, so hipace.dt = adaptive works across ranks: every rank keeps
// its own dt and min_uz_mq exactly as the reference's ranks do); plasma is re-created locally.  The laser envelope slices
// A^{n+1}, A^n of a slice ride in the same group of sends / receives (MultiBuffer.cpp:444-490).
//
// Each directed edge r -> r+1 is its own 2-rank communicator with its own CUDA stream, so a
// receive that waits for the upstream rank never blocks the sends to the downstream rank.
// Receives are posted `lookahead` slices ahead of the slice loop into the beam ring (one
// fixed-capacity packet per slice, sim.hpp), the compute stream waits on a per-slice event; sends
// are enqueued behind a per-slice event of the compute stream.  No host synchronisation and no
// particle count on the host: packets carry their counts in their header.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already loaded, else the
// system one) so that libhpb200.so has no link-time dependency on it.
#include "sim.hpp"
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

namespace {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi *nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.h ? &api : nullptr;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.h) break;
    }
    if (!api.h) { hpb_set_error("pipeline: cannot load libnccl.so.2: %s", dlerror()); return nullptr; }
#define HPB_SYM(field, name)                                                              \
    *(void **)(&api.field) = dlsym(api.h, name);                                          \
    if (!api.field) { hpb_set_error("pipeline: libnccl lacks %s", name); api.h = nullptr; return nullptr; }
    HPB_SYM(GetUniqueId, "ncclGetUniqueId")
    HPB_SYM(CommInitRank, "ncclCommInitRank")
    HPB_SYM(CommDestroy, "ncclCommDestroy")
    HPB_SYM(Send, "ncclSend")
    HPB_SYM(Recv, "ncclRecv")
    HPB_SYM(GroupStart, "ncclGroupStart")
    HPB_SYM(GroupEnd, "ncclGroupEnd")
    HPB_SYM(GetErrorString, "ncclGetErrorString")
#undef HPB_SYM
    return &api;
}

#define HPB_NCCL(expr)                                                                       \
    do {                                                                                     \
        ncclResult_t r_ = (expr);                                                            \
        if (r_ != ncclSuccess) {                                                             \
            hpb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, N->GetErrorString(r_)); \
            return HPB_ERR_NCCL;                                                             \
        }                                                                                    \
    } while (0)

}  // namespace

struct hpb_pipeline {
    int rank = 0, world = 1;
    ncclComm_t comm_recv = nullptr, comm_send = nullptr;   // 2-rank comms: sender is rank 0, receiver rank 1
    cudaStream_t s_recv = nullptr, s_send = nullptr;
    std::vector<cudaEvent_t> ev_recv, ev_ready, ev_sent;   // per slot
    std::vector<char> sent_pending;                        // per slot: ev_sent[slot] guards the out ring
    cudaEvent_t ev_step_done = nullptr;
    bool step_done_recorded = false;
    bool receiving = false, sent_this_step = false;
    int posted = 0;                                        // slots with a receive posted this step
    int lookahead = 8;
    // the physical time of the next step (MultiBuffer::put_time / get_time): a ring of send slots (a slot
    // is reused 8 owned steps later, long after its send left) and one receive slot
    double *d_time = nullptr, *h_time = nullptr;           // [0..7] send, [8] receive; h_time pinned
    unsigned n_time_sent = 0;
};

bool hpb_pipeline_active(const hpb_sim *s) { return s->pipe && s->pipe->world > 1; }
int hpb_pipeline_world(const hpb_sim *s) { return s->pipe ? s->pipe->world : 1; }
int hpb_pipeline_rank(const hpb_sim *s) { return s->pipe ? s->pipe->rank : 0; }

// every time step but the first has an upstream (Hipace.cpp:410: step 0 starts from the initial
// beam, later steps from MultiBuffer::get_data)
bool hpb_pipeline_receives(const hpb_sim *s, int step) { return hpb_pipeline_active(s) && step > 0; }

extern "C" int hpb_nccl_unique_id(char out[HPB_NCCL_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == HPB_NCCL_ID_BYTES, "ncclUniqueId size");
    NcclApi *N = nccl();
    if (!N || !out) return HPB_ERR_NCCL;
    ncclUniqueId id;
    HPB_NCCL(N->GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return HPB_OK;
}

extern "C" int hpb_sim_pipeline_init(hpb_sim *s, int rank, int world, const char *id_recv,
                                     const char *id_send)
{
    if (!s || world < 1 || rank < 0 || rank >= world) return HPB_ERR_ARG;
    SIM_CUDA(cudaSetDevice(s->device));
    hpb_pipeline_destroy(s);
    hpb_pipeline *p = new hpb_pipeline();
    p->rank = rank; p->world = world;
    s->pipe = p;
    if (world == 1) return HPB_OK;
    if (!id_recv || !id_send) return HPB_ERR_ARG;
    // one channel (= one CTA) per point-to-point kernel unless the user decided otherwise: the packets
    // are small and the CTAs of a waiting receive otherwise occupy many SMs (takes effect only if NCCL
    // has not read its parameters yet; bench.py and hipace_b200.pipeline set it before torch starts NCCL)
    setenv("NCCL_MAX_P2P_NCHANNELS", "1", 0);
    setenv("NCCL_MIN_P2P_NCHANNELS", "1", 0);
    NcclApi *N = nccl();
    if (!N) return HPB_ERR_NCCL;
    ncclUniqueId idr, ids;
    memcpy(&idr, id_recv, sizeof(idr));
    memcpy(&ids, id_send, sizeof(ids));
    // edges are created in increasing edge index (edge e: rank e -> rank e+1 mod R): a consistent
    // global order, so the blocking 2-rank initialisations cannot dead-lock
    const int e_send = rank, e_recv = (rank - 1 + world) % world;
    for (int pass = 0; pass < 2; ++pass) {
        const bool do_send = (pass == 0) == (e_send < e_recv);
        if (do_send) HPB_NCCL(N->CommInitRank(&p->comm_send, 2, ids, 0));
        else HPB_NCCL(N->CommInitRank(&p->comm_recv, 2, idr, 1));
    }
    SIM_CUDA(cudaStreamCreateWithFlags(&p->s_recv, cudaStreamNonBlocking));
    SIM_CUDA(cudaStreamCreateWithFlags(&p->s_send, cudaStreamNonBlocking));
    p->ev_recv.resize(s->nz); p->ev_ready.resize(s->nz); p->ev_sent.resize(s->nz);
    p->sent_pending.assign(s->nz, 0);
    for (auto &e : p->ev_recv) SIM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : p->ev_ready) SIM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto &e : p->ev_sent) SIM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    SIM_CUDA(cudaEventCreateWithFlags(&p->ev_step_done, cudaEventDisableTiming));
    SIM_CUDA(cudaMalloc(&p->d_time, 9 * sizeof(double)));
    SIM_CUDA(cudaMallocHost(&p->h_time, 9 * sizeof(double)));
    // connect both edges now: NCCL sets up its P2P channels on the first send / receive of a
    // communicator, and that set-up BLOCKS THE HOST until the peer enters the matching call.  As
    // for the communicators themselves the two edges of a rank are therefore connected in
    // increasing edge index (rank 0: send on edge 0, then receive on edge R-1; every other rank:
    // receive on edge r-1, then send on edge r) -- a chain, no cycle.
    char *d_hs = nullptr;
    SIM_CUDA(cudaMalloc(&d_hs, 512));
    SIM_CUDA(cudaMemset(d_hs, 0, 512));
    for (int pass = 0; pass < 2; ++pass) {
        const bool do_send = (pass == 0) == (e_send < e_recv);
        if (do_send) {
            HPB_NCCL(N->Send(d_hs, 256, ncclUint8, 1, p->comm_send, p->s_send));
            SIM_CUDA(cudaStreamSynchronize(p->s_send));
        } else {
            HPB_NCCL(N->Recv(d_hs + 256, 256, ncclUint8, 0, p->comm_recv, p->s_recv));
            SIM_CUDA(cudaStreamSynchronize(p->s_recv));
        }
    }
    cudaFree(d_hs);
    return HPB_OK;
}

void hpb_pipeline_destroy(hpb_sim *s)
{
    hpb_pipeline *p = s->pipe;
    if (!p) return;
    NcclApi *N = nccl();
    if (p->s_recv) cudaStreamSynchronize(p->s_recv);
    if (p->s_send) cudaStreamSynchronize(p->s_send);
    // ncclCommDestroy may wait for the peer of the communicator: tear the two edges down in the
    // same global order (increasing edge index) in which they were built
    const int e_send = p->rank, e_recv = (p->rank - 1 + p->world) % p->world;
    for (int pass = 0; pass < 2; ++pass) {
        const bool do_send = (pass == 0) == (e_send < e_recv);
        if (do_send) { if (N && p->comm_send) N->CommDestroy(p->comm_send); }
        else if (N && p->comm_recv) N->CommDestroy(p->comm_recv);
    }
    for (auto &e : p->ev_recv) cudaEventDestroy(e);
    for (auto &e : p->ev_ready) cudaEventDestroy(e);
    for (auto &e : p->ev_sent) cudaEventDestroy(e);
    if (p->ev_step_done) cudaEventDestroy(p->ev_step_done);
    cudaFree(p->d_time);
    if (p->h_time) cudaFreeHost(p->h_time);
    if (p->s_recv) cudaStreamDestroy(p->s_recv);
    if (p->s_send) cudaStreamDestroy(p->s_send);
    delete p;
    s->pipe = nullptr;
}

extern "C" long hpb_sim_pipeline_message_bytes(hpb_sim *s)
{
    if (!s) return -1;
    long n = 0;
    for (auto &b : s->beams) n += (long)b.ring[0].msg_bytes();
    return n;
}

static int post_receives(hpb_sim *s, int upto)
{
    hpb_pipeline *p = s->pipe;
    NcclApi *N = nccl();
    if (upto > s->nz) upto = s->nz;
    for (; p->posted < upto; ++p->posted) {
        const int slot = p->posted;
        HPB_NCCL(N->GroupStart());
        for (auto &b : s->beams) {
            const BeamRing &in = b.ring[b.cur];
            HPB_NCCL(N->Recv(in.packet(slot), in.msg_bytes(), ncclUint8, 0, p->comm_recv, p->s_recv));
        }
        if (s->laser_state) {       // A^n, A^{n-1} of this slice (MultiBuffer.cpp:444-490)
            void *rv[2], *sd[2]; size_t nb = 0;
            hpb_laser_packet(s->laser_state, s->nz - 1 - slot, rv, sd, &nb);
            HPB_NCCL(N->Recv(rv[0], nb, ncclUint8, 0, p->comm_recv, p->s_recv));
            HPB_NCCL(N->Recv(rv[1], nb, ncclUint8, 0, p->comm_recv, p->s_recv));
        }
        HPB_NCCL(N->GroupEnd());
        SIM_CUDA(cudaEventRecord(p->ev_recv[slot], p->s_recv));
    }
    return HPB_OK;
}

// MultiBuffer::put_time (utils/MultiBuffer.cpp:629-651): the time at which the NEXT step starts leaves for
// the downstream rank at the beginning of this step -- one 8-byte message per step on the edge's own
// stream, ahead of the step's slice packets
int hpb_pipeline_put_time(hpb_sim *s, int step, double t_next)
{
    hpb_pipeline *p = s->pipe;
    if (!hpb_pipeline_active(s) || step + 1 > s->max_step) return HPB_OK;     // Hipace.cpp:445-447
    NcclApi *N = nccl();
    const unsigned k = p->n_time_sent++ % 8;
    p->h_time[k] = t_next;
    SIM_CUDA(cudaMemcpyAsync(p->d_time + k, p->h_time + k, sizeof(double), cudaMemcpyHostToDevice, p->s_send));
    HPB_NCCL(N->Send(p->d_time + k, sizeof(double), ncclUint8, 1, p->comm_send, p->s_send));
    return HPB_OK;
}

// MultiBuffer::get_time (:611-627): blocks the host like the reference's MPI_Recv; the upstream rank posted
// the value when it began the step before this one
int hpb_pipeline_get_time(hpb_sim *s, int step, double *t)
{
    hpb_pipeline *p = s->pipe;
    if (!hpb_pipeline_receives(s, step)) return HPB_OK;
    NcclApi *N = nccl();
    HPB_NCCL(N->Recv(p->d_time + 8, sizeof(double), ncclUint8, 0, p->comm_recv, p->s_recv));
    SIM_CUDA(cudaMemcpyAsync(p->h_time + 8, p->d_time + 8, sizeof(double), cudaMemcpyDeviceToHost, p->s_recv));
    SIM_CUDA(cudaStreamSynchronize(p->s_recv));
    *t = p->h_time[8];
    return HPB_OK;
}

int hpb_pipeline_begin_step(hpb_sim *s, int step)
{
    hpb_pipeline *p = s->pipe;
    p->receiving = hpb_pipeline_receives(s, step);
    p->sent_this_step = false;
    p->posted = 0;
    if (!p->receiving) return HPB_OK;
    // We receive into the ring the previous owned step used as its INPUT (begin_step flips
    // BeamSp::cur back when the other ring still holds slices being sent): it is free as soon as
    // that step's slice loop is done.  The ring with the pending sends becomes this step's output
    // ring; every slot of it is guarded by its own send-complete event (hpb_pipeline_wait_out_slot),
    // so a rank never waits for the downstream rank to drain a whole step -- with a single
    // "all sent" event two ranks ran their steps one after the other instead of pipelined.
    if (p->step_done_recorded) SIM_CUDA(cudaStreamWaitEvent(p->s_recv, p->ev_step_done, 0));
    return post_receives(s, p->lookahead);
}

int hpb_pipeline_wait_slice_on(hpb_sim *s, int islice, cudaStream_t st)
{
    hpb_pipeline *p = s->pipe;
    if (!p || !p->receiving) return HPB_OK;
    const int slot = s->nz - 1 - islice;
    int rc = post_receives(s, slot + 1 + p->lookahead);
    if (rc) return rc;
    SIM_CUDA(cudaStreamWaitEvent(st, p->ev_recv[slot], 0));
    return HPB_OK;
}
int hpb_pipeline_wait_slice(hpb_sim *s, int islice)
{
    return hpb_pipeline_wait_slice_on(s, islice, s->beam_stream ? s->beam_stream : s->stream);
}

int hpb_pipeline_send_slice(hpb_sim *s, int islice, int step)
{
    hpb_pipeline *p = s->pipe;
    if (!hpb_pipeline_active(s) || step + 1 > s->max_step) return HPB_OK;     // Hipace.cpp:441-443
    NcclApi *N = nccl();
    const int slot = s->nz - 1 - islice;
    SIM_CUDA(cudaEventRecord(p->ev_ready[slot], s->beam_stream ? s->beam_stream : s->stream));
    SIM_CUDA(cudaStreamWaitEvent(p->s_send, p->ev_ready[slot], 0));
    HPB_NCCL(N->GroupStart());
    for (auto &b : s->beams) {
        const BeamRing &out = b.ring[b.cur ^ 1];
        HPB_NCCL(N->Send(out.packet(slot), out.msg_bytes(), ncclUint8, 1, p->comm_send, p->s_send));
    }
    if (s->laser_state) {           // A^{n+1}, A^n of this slice for the next time step
        void *rv[2], *sd[2]; size_t nb = 0;
        hpb_laser_packet(s->laser_state, islice, rv, sd, &nb);
        HPB_NCCL(N->Send(sd[0], nb, ncclUint8, 1, p->comm_send, p->s_send));
        HPB_NCCL(N->Send(sd[1], nb, ncclUint8, 1, p->comm_send, p->s_send));
    }
    HPB_NCCL(N->GroupEnd());
    SIM_CUDA(cudaEventRecord(p->ev_sent[slot], p->s_send));
    p->sent_pending[slot] = 1;
    p->sent_this_step = true;
    return HPB_OK;
}

// the compute stream is about to overwrite slot `slot` of the output ring: the send of the slice
// it held (previous owned step) must have left the buffer
int hpb_pipeline_wait_out_slot_on(hpb_sim *s, int islice, cudaStream_t st)
{
    hpb_pipeline *p = s->pipe;
    if (!hpb_pipeline_active(s)) return HPB_OK;
    const int slot = s->nz - 1 - islice;
    if (p->sent_pending[slot]) {
        SIM_CUDA(cudaStreamWaitEvent(st, p->ev_sent[slot], 0));
        p->sent_pending[slot] = 0;
    }
    return HPB_OK;
}
int hpb_pipeline_wait_out_slot(hpb_sim *s, int islice)
{
    return hpb_pipeline_wait_out_slot_on(s, islice, s->beam_stream ? s->beam_stream : s->stream);
}

bool hpb_pipeline_out_ring_busy(const hpb_sim *s)
{
    if (!hpb_pipeline_active(s)) return false;
    for (char c : s->pipe->sent_pending) if (c) return true;
    return false;
}

int hpb_pipeline_end_step(hpb_sim *s, int step)
{
    (void)step;
    hpb_pipeline *p = s->pipe;
    SIM_CUDA(cudaEventRecord(p->ev_step_done, s->stream));
    p->step_done_recorded = true;
    return HPB_OK;
}
