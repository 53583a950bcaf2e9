// The ONE exchange of the sharded path: all-reduce(sum) of the packed fp64 statistics vector
// [loss | class histogram(C) | confusion matrix(C*C)] (8(1+C+C*C) bytes, exact below 2^53) over NCCL / NVLink.
//
// torch.distributed's ProcessGroupNCCL costs ~25 us of host time per collective, which at the
// reference's batch sizes (a 35 us step) is most of a step; here the collective is enqueued by this
// library straight on an NCCL communicator of its own: the unique id is created on rank 0 and
// handed to the other ranks by the caller (torch.distributed broadcast -- plumbing), and each
// all-reduce is one ncclAllReduce on a side stream forked from the caller's stream after the
// finalisation kernel, so that it overlaps the backward kernel; the caller's stream joins it only
// when the statistics are consumed.  libnccl.so.2 is resolved at run time (dlopen: PyTorch has it
// loaded already), so the library has no link-time NCCL dependency.
#include <dlfcn.h>
#include <new>
#include "common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(nccl_comm_t*, int, nccl_unique_id, int);
typedef int (*fn_comm_destroy)(nccl_comm_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
    fn_get_unique_id get_unique_id;
    fn_comm_init_rank comm_init_rank;
    fn_comm_destroy comm_destroy;
    fn_all_reduce all_reduce;
    bool ok;
};

const NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a = {};
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            a.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
            a.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
            a.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
            a.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
            a.ok = a.get_unique_id && a.comm_init_rank && a.comm_destroy && a.all_reduce;
        }
        return a;
    }();
    return api;
}

constexpr int kNcclFloat64 = 8, kNcclSum = 0;      // nccl.h: ncclDataType_t / ncclRedOp_t
constexpr int kRing = 8;

}  // namespace

struct msq_comm {
    nccl_comm_t comm;
    cudaStream_t side;
    cudaEvent_t fork, done[kRing];        // done[n % kRing]: completion of the n-th all-reduce
    int world, rank;
    unsigned long long issued;
    // peer-memory mailboxes (see PeerBox, common.cuh): used by msq_fused_fwd_bwd instead of ncclAllReduce
    uint4* box_local;                     // this rank's mailbox (cudaMalloc)
    uint4* box_peer[msq::kMaxPeers];      // every rank's mailbox as mapped here (box_peer[rank] == box_local)
    unsigned* box_err;                    // device error word
    msq::PeerBoxStatic* box_static;       // device copy of {world, rank, err, peers}
    bool box_mapped, box_ready;           // peers mapped / mailbox path switched on (msq_comm_box_enable)
    unsigned box_seq;                     // steps issued so far = sequence number of the newest vector
    double* box_last;                     // statistics of step box_seq: produced, not pushed yet (NULL: none pending)
    double* box_pushed;                   // statistics of step box_seq-1: pushed, not reduced yet (NULL: none pending)
    int box_last_count, box_pushed_count;
};

namespace {
constexpr size_t kBoxBytes = sizeof(uint4) * msq::kBoxSlots * msq::kMaxPeers * msq::kBoxCount;

// completes the two steps still in flight (the steps themselves carry the exchange in their finalisation kernels):
// pushes the newest vector, reduces the one before it, then reduces the newest in place
__global__ void __launch_bounds__(32) box_flush_kernel(const msq::PeerBox box, double* last) {
    msq::box_exchange(box.st, box.cur, box.prev_out, box.seq, box.count, box.prev_count, (int)threadIdx.x);
    __syncwarp();
    msq::box_reduce(box.st, box.seq, box.count, last, (int)threadIdx.x);
}

msq::PeerBox make_box(const msq_comm* c) {
    msq::PeerBox b = {};
    b.st = c->box_static;
    return b;
}
}  // namespace

extern "C" int msq_comm_unique_id(void* id128) {
    if (!id128) return MSQ_E_BADARG;
    if (!nccl().ok) return MSQ_E_NCCL;
    return nccl().get_unique_id((nccl_unique_id*)id128) == 0 ? 0 : MSQ_E_NCCL;
}

extern "C" int msq_comm_create(const void* id128, int world, int rank, msq_comm** out) {
    if (!id128 || !out || world < 1 || rank < 0 || rank >= world) return MSQ_E_BADARG;
    if (!nccl().ok) return MSQ_E_NCCL;
    msq_comm* c = new (std::nothrow) msq_comm();
    if (!c) return (int)cudaErrorMemoryAllocation;
    c->world = world; c->rank = rank; c->issued = 0;
    c->box_local = nullptr; c->box_err = nullptr; c->box_static = nullptr; c->box_mapped = false; c->box_ready = false; c->box_seq = 0u; c->box_last = nullptr; c->box_pushed = nullptr; c->box_last_count = 0; c->box_pushed_count = 0;
    for (int p = 0; p < msq::kMaxPeers; ++p) c->box_peer[p] = nullptr;
    nccl_unique_id id;
    memcpy(&id, id128, sizeof(id));
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return (int)e; }
    cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming);
    for (int i = 0; i < kRing; ++i) cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
    if (nccl().comm_init_rank(&c->comm, world, id, rank) != 0) {
        cudaEventDestroy(c->fork);
        for (int i = 0; i < kRing; ++i) cudaEventDestroy(c->done[i]);
        cudaStreamDestroy(c->side);
        delete c;
        return MSQ_E_NCCL;
    }
    *out = c;
    return 0;
}

// Enqueue all-reduce(sum) of buf[0..count) (device fp64, in place) after everything already enqueued on
// `stream`; returns at once.  The result may be read by work enqueued on `stream` after msq_comm_join.
extern "C" int msq_comm_allreduce_f64(msq_comm* c, double* buf, int count, msq_stream_t stream) {
    if (!c || !buf || count < 1) return MSQ_E_BADARG;
    cudaError_t e;
    if ((e = cudaEventRecord(c->fork, (cudaStream_t)stream)) != cudaSuccess) return (int)e;
    if ((e = cudaStreamWaitEvent(c->side, c->fork, 0)) != cudaSuccess) return (int)e;
    if (nccl().all_reduce(buf, buf, (size_t)count, kNcclFloat64, kNcclSum, c->comm, c->side) != 0) return MSQ_E_NCCL;
    if ((e = cudaEventRecord(c->done[c->issued % kRing], c->side)) != cudaSuccess) return (int)e;
    c->issued++;
    return 0;
}

// Make `stream` wait for the all-reduce issued `lag` calls before the most recent one (lag 0 = the most
// recent; no host synchronisation).  A lag of 1-2 lets a collective take more than one step without stalling
// the kernels: the statistics it carries are only logged.
extern "C" int msq_comm_join(msq_comm* c, int lag, msq_stream_t stream) {
    if (!c || lag < 0 || lag >= kRing) return MSQ_E_BADARG;
    if (c->box_ready && c->box_last && lag == 0) {
        // peer-memory path: the newest step's vector is not pushed yet, the one before it not reduced yet
        msq::PeerBox b = make_box(c);
        b.cur = c->box_last;
        b.seq = c->box_seq;
        b.count = (short)c->box_last_count;
        b.prev_out = c->box_pushed;
        b.prev_count = (short)c->box_pushed_count;
        box_flush_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(b, c->box_last);
        c->box_last = nullptr;
        c->box_pushed = nullptr;
        MSQ_CHECK_LAUNCH();
    }
    if (c->issued <= (unsigned long long)lag) return 0;
    const cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, c->done[(c->issued - 1 - lag) % kRing], 0);
    return (int)e;
}

// ---- peer-memory mailboxes -------------------------------------------------------------------------------
// msq_comm_box_export: allocate and zero this rank's mailbox, return its 64-byte cudaIpc handle.  The caller
// all-gathers the handles (torch.distributed: plumbing) and hands all of them, in rank order, to msq_comm_box_open,
// which maps the peers' mailboxes (NVLink peer access).  Both are collective over the ranks of the communicator; if
// either fails on any rank the caller simply does not use the mailboxes (msq_fused_fwd_bwd then all-reduces with NCCL).
extern "C" int msq_comm_box_export(msq_comm* c, void* handle64) {
    if (!c || !handle64) return MSQ_E_BADARG;
    if (c->world > msq::kMaxPeers) return MSQ_E_BADARG;
    cudaError_t e;
    if (!c->box_local) {
        constexpr size_t kTail = 16 + sizeof(msq::PeerBoxStatic);
        if ((e = cudaMalloc((void**)&c->box_local, kBoxBytes + kTail)) != cudaSuccess) return (int)e;
        if ((e = cudaMemset(c->box_local, 0, kBoxBytes + kTail)) != cudaSuccess) return (int)e;
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) return (int)e;
        c->box_err = (unsigned*)((char*)c->box_local + kBoxBytes);
        c->box_static = (msq::PeerBoxStatic*)((char*)c->box_local + kBoxBytes + 16);
    }
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, c->box_local)) != cudaSuccess) return (int)e;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int msq_comm_box_open(msq_comm* c, const void* handles /* world x 64 bytes, rank order */) {
    if (!c || !handles || !c->box_local || c->world > msq::kMaxPeers) return MSQ_E_BADARG;
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) { c->box_peer[p] = c->box_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * p, 64);
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
        c->box_peer[p] = (uint4*)ptr;
    }
    msq::PeerBoxStatic st = {};
    st.world = c->world;
    st.rank = c->rank;
    st.err = c->box_err;
    for (int p = 0; p < c->world; ++p) st.peer[p] = c->box_peer[p];
    cudaError_t e = cudaMemcpy(c->box_static, &st, sizeof(st), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return (int)e;
    c->box_mapped = true;
    return 0;
}

// Switch the mailbox path on (only if every peer is mapped) or off.  The caller enables it on all ranks or on none:
// after msq_comm_box_open it agrees on the outcome across the ranks (an all-reduce(min) of "my open succeeded").
extern "C" int msq_comm_box_enable(msq_comm* c, int on) {
    if (!c) return MSQ_E_BADARG;
    if (on && !c->box_mapped) return MSQ_E_BADARG;
    if (c->box_last) return MSQ_E_BADARG;          // steps in flight: msq_comm_join first
    c->box_ready = on != 0;
    return 0;
}

// 1 if msq_fused_fwd_bwd exchanges the statistics through the mailboxes, 0 if it uses ncclAllReduce
extern "C" int msq_comm_box_active(const msq_comm* c) { return (c && c->box_ready) ? 1 : 0; }

// device error word of the mailbox path (bit 0: a peer's vector never arrived); synchronises the device
extern "C" int msq_comm_box_errors(msq_comm* c, unsigned* out) {
    if (!c || !out) return MSQ_E_BADARG;
    *out = 0u;
    if (!c->box_err) return 0;
    return (int)cudaMemcpy(out, c->box_err, sizeof(unsigned), cudaMemcpyDeviceToHost);
}

extern "C" void msq_comm_destroy(msq_comm* c) {
    if (!c) return;
    cudaStreamSynchronize(c->side);
    if (c->box_local) {
        cudaDeviceSynchronize();
        for (int p = 0; p < c->world && p < msq::kMaxPeers; ++p)
            if (p != c->rank && c->box_peer[p]) cudaIpcCloseMemHandle(c->box_peer[p]);
        cudaFree(c->box_local);
    }
    if (c->comm) nccl().comm_destroy(c->comm);
    cudaEventDestroy(c->fork);
    for (int i = 0; i < kRing; ++i) cudaEventDestroy(c->done[i]);
    cudaStreamDestroy(c->side);
    delete c;
}

// One call per training step for callers that know the upstream gradient scale when they call the forward
// (lambda_target is a constant, tools/solve_gta5.py:199,217): msq_fused_fwd + msq_fused_bwd and, when the images are
// sharded over ranks (comm != NULL), the exchange of the step's statistics vector: carried by the finalisation kernel
// over the peer-memory mailboxes when they are open, else one ncclAllReduce forked after the backward (so that nothing
// sits between forward -> finalise -> backward) and ordered after the collective issued `lag` steps earlier.  Same
// kernels and results as the separate calls; it also keeps the host side of a 35 us step to one library call.
// All steps of one communicator must be enqueued on the same stream (the mailbox protocol relies on stream order).
extern "C" int msq_fused_fwd_bwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                                 double ratio, int n_images_norm, void* accum, void* out, void* aux, const float* grad_out,
                                 float grad_scale, float* grad_logits, msq_comm* comm, int lag, msq_stream_t stream) {
    if (!grad_logits || !out || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES) return MSQ_E_BADARG;
    if ((((uintptr_t)aux) & 15u) || (((uintptr_t)grad_logits) & 3u)) return MSQ_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    const msq_state_layout lay = msq::make_layout(n, num_class);
    double* stats = (double*)((char*)out + lay.stats_off);
    if (comm && comm->box_ready && 1 + num_class <= msq::kBoxCount) {
        // sharded: a second CTA of this step's finalisation kernel pushes the PREVIOUS step's [loss | hist] into every
        // rank's mailbox over NVLink and reduces the step before that; nothing is enqueued between or after the three
        // kernels of the step and the hot kernels are untouched
        msq::PeerBox b = {};
        if (comm->box_last) {
            b = make_box(comm);
            b.cur = comm->box_last;
            b.seq = comm->box_seq;
            b.count = (short)comm->box_last_count;
            b.prev_out = comm->box_pushed;
            b.prev_count = (short)comm->box_pushed_count;
        }
        int rc = msq::fused_fwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, nullptr, ratio, n_images_norm, accum,
                                         out, aux, grad_logits, s, 0, &b);
        if (rc) return rc;
        comm->box_pushed = comm->box_last;
        comm->box_pushed_count = comm->box_last_count;
        comm->box_last = stats;
        comm->box_last_count = 1 + num_class;
        comm->box_seq += 1u;
        return msq::fused_bwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, n_images_norm, out, grad_out, grad_scale,
                                       grad_logits, aux, 1, s);
    }
    int rc = msq::fused_fwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, nullptr, ratio, n_images_norm, accum, out,
                                     aux, grad_logits, s);
    if (rc) return rc;
    rc = msq::fused_bwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, n_images_norm, out, grad_out, grad_scale,
                                 grad_logits, aux, 1, s);
    if (rc || !comm) return rc;
    rc = msq_comm_join(comm, lag, stream);
    if (rc) return rc;
    return msq_comm_allreduce_f64(comm, stats, 1 + num_class, stream);
}
