// The ONE exchange of the sharded path: all-reduce(sum) of the packed fp64 statistics vector
// [loss | class histogram(C) | confusion matrix(C*C)] (8(1+C+C*C) bytes, exact below 2^53) over NCCL / NVLink.
//
// torch.distributed's ProcessGroupNCCL costs ~25 us of host time per collective, which at the
// reference's batch sizes (a 35 us step) is most of a step; here the collective is enqueued by this
// library straight on an NCCL communicator of its own: the unique id is created on rank 0 and
// handed to the other ranks by the caller (torch.distributed broadcast -- plumbing), and each
// all-reduce is one ncclAllReduce on a side stream forked from the caller's stream, so that it
// overlaps whatever the caller enqueues next; the caller's stream joins it only when the statistics
// are consumed.  libnccl.so.2 is resolved at run time (dlopen: PyTorch has it
// loaded already), so the library has no link-time NCCL dependency.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
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

constexpr int kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0;      // nccl.h: ncclDataType_t / ncclRedOp_t
constexpr int kRing = 8;

}  // namespace

struct msq_comm {
    nccl_comm_t comm;
    cudaStream_t side;
    cudaEvent_t fork, done[kRing];        // done[n % kRing]: completion of the n-th all-reduce
    int world, rank;
    unsigned long long issued;
    // ---- per-step statistics of msq_fused_fwd_bwd.  Step k (1, 2, ...) owns slot k % kBoxSlots of two device rings the
    //      COMMUNICATOR allocates: no pointer into a caller's buffer is kept across calls (round-1 advice) ----
    double* ring;                         // cudaMalloc: [2][kBoxSlots][kBoxCount]
    double* box_vec;                      // ring + 0: this rank's vector of step k (written by that step's finalisation)
    double* box_red;                      // ring + 1: the all-reduced vector of step k
    int count[msq::kBoxSlots];            // doubles in the vector of the step that owns the slot
    unsigned long long ar_index[msq::kBoxSlots];      // NCCL path: which all-reduce (c->issued numbering) carries the slot
    unsigned seq;                         // steps issued
    unsigned pushed, reduced;             // mailbox path: highest step whose push / reduction has been enqueued
    // ---- peer-memory mailboxes (see PeerBox, common.cuh) ----
    uint4* box_local;                     // this rank's mailbox (cudaMalloc)
    uint4* box_peer[msq::kMaxPeers];      // every rank's mailbox as mapped here (box_peer[rank] == box_local)
    unsigned* err_host;                   // cudaHostAlloc(mapped): [0] bits, [1] vectors lost so far
    unsigned err_reported;                // losses already returned to the caller as MSQ_E_PEER
    msq::PeerBoxStatic* box_static;       // device copy of {world, rank, err, timeout, peers}
    msq::PeerBoxStatic box_static_host;
    bool box_mapped, box_ready;           // peers mapped / mailbox path switched on (msq_comm_box_enable)
    // ---- same-step sum of a few uint64 words (msq_comm_sum_u64_begin / _end) ----
    unsigned u64_seq;                     // exchanges begun (mailbox cells kBoxCount-2.. carry them: flags count independently)
    int u64_pending;                      // words of the exchange begun and not yet ended (0: none)
    unsigned long long* u64_buf;          // NCCL path: [kU64Ring][kU64Max] staging words (cudaMalloc)
};

namespace {
constexpr size_t kBoxBytes = sizeof(uint4) * msq::kBoxSlots * msq::kMaxPeers * msq::kBoxCount;
constexpr size_t kRingDoubles = (size_t)msq::kBoxSlots * msq::kBoxCount;

// completes the steps still in flight (the steps themselves carry the exchange in an extra CTA of their own kernels):
// pushes the newest vector and reduces the one before it (box), then reduces the newest (last_seq -> last_out)
__global__ void __launch_bounds__(32) box_flush_kernel(const msq::PeerBox box, unsigned last_seq, int last_count, double* last_out) {
    msq::box_exchange(box, (int)threadIdx.x);
    __syncwarp();
    if (last_out) msq::box_reduce(box.st, last_seq, last_count, last_out, (int)threadIdx.x);
}

constexpr int kU64Max = 2, kU64Ring = 4;
constexpr int kU64Cell = msq::kBoxCount - kU64Max;       // mailbox cells 38, 39: never used by a statistics vector (<= 33 doubles)

// Same-step exchange over the mailboxes, split in two tiny kernels so that the caller can put independent work between them:
// push this rank's words into every rank's ring ...
__global__ void __launch_bounds__(32) u64_push_kernel(const msq::PeerBoxStatic* st, unsigned seq, const unsigned long long* src, int count) {
    const int lane = (int)threadIdx.x;
    if (lane >= count) return;
    const double v = __longlong_as_double((long long)src[lane]);          // the cell carries 64 raw bits
    for (int p = 0; p < st->world; ++p) msq::ll_store(msq::box_cell(st->peer[p], seq, st->rank, kU64Cell + lane), v, seq);
}
// ... and sum, as INTEGERS, what all ranks pushed for the same sequence number (spins until they have; time-out as box_reduce)
__global__ void __launch_bounds__(32) u64_reduce_kernel(const msq::PeerBoxStatic* st, unsigned seq, unsigned long long* dst, int count) {
    const int lane = (int)threadIdx.x;
    uint4* mine = st->peer[st->rank];
    const unsigned long long limit = st->timeout_ns;
    unsigned long long t0 = 0ull, sum = 0ull;
    bool lost = false;
    if (lane < count) {
        for (int p = 0; p < st->world; ++p) {
            const uint4* cell = msq::box_cell(mine, seq, p, kU64Cell + lane);
            double v = 0.0;
            unsigned spins = 0;
            while (!lost && !msq::ll_load(cell, seq, v)) {
                if ((++spins & 255u) == 0u) {
                    const unsigned long long now = msq::global_ns();
                    if (t0 == 0ull) t0 = now;
                    else if (now - t0 > limit) lost = true;
                }
                __nanosleep(64);
            }
            sum += lost ? 0ull : (unsigned long long)__double_as_longlong(v);
        }
        dst[lane] = lost ? 0ull : sum;                                    // a lost peer: 0 (a zero count divides to NaN downstream)
    }
    if (__any_sync(0xffffffffu, lost) && lane == 0) {
        volatile unsigned* e = st->err;
        e[0] = e[0] | 1u;
        e[1] = e[1] + 1u;
    }
}

inline double* vec_slot(const msq_comm* c, unsigned seq) { return c->box_vec + (size_t)(seq % msq::kBoxSlots) * msq::kBoxCount; }
inline double* red_slot(const msq_comm* c, unsigned seq) { return c->box_red + (size_t)(seq % msq::kBoxSlots) * msq::kBoxCount; }

// a peer's vector was lost since the last call: reported ONCE per loss (stale by at most the steps in flight: the words
// live in mapped host memory and are read without synchronising)
int peer_error(msq_comm* c) {
    if (!c->err_host) return 0;
    const unsigned lost = ((volatile unsigned*)c->err_host)[1];
    if (lost != c->err_reported) { c->err_reported = lost; return MSQ_E_PEER; }
    return 0;
}

unsigned long long default_timeout_ns() {
    const char* e = getenv("MSQ_BOX_TIMEOUT_S");
    double sec = e ? atof(e) : 600.0;
    if (!(sec > 0.0)) sec = 600.0;
    return (unsigned long long)(sec * 1e9);
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
    msq_comm* c = new (std::nothrow) msq_comm();          // value-initialised: every member zero
    if (!c) return (int)cudaErrorMemoryAllocation;
    c->world = world; c->rank = rank;
    nccl_unique_id id;
    memcpy(&id, id128, sizeof(id));
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking)) != cudaSuccess) { delete c; return (int)e; }
    if ((e = cudaMalloc((void**)&c->ring, 2 * kRingDoubles * sizeof(double))) != cudaSuccess) { cudaStreamDestroy(c->side); delete c; return (int)e; }
    cudaMemset(c->ring, 0, 2 * kRingDoubles * sizeof(double));
    c->box_vec = c->ring;
    c->box_red = c->ring + kRingDoubles;
    if ((e = cudaMalloc((void**)&c->u64_buf, sizeof(unsigned long long) * kU64Ring * kU64Max)) != cudaSuccess) {
        cudaStreamDestroy(c->side); cudaFree(c->ring); delete c; return (int)e;
    }
    cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming);
    for (int i = 0; i < kRing; ++i) cudaEventCreateWithFlags(&c->done[i], cudaEventDisableTiming);
    if (nccl().comm_init_rank(&c->comm, world, id, rank) != 0) {
        cudaEventDestroy(c->fork);
        for (int i = 0; i < kRing; ++i) cudaEventDestroy(c->done[i]);
        cudaStreamDestroy(c->side);
        cudaFree(c->ring);
        delete c;
        return MSQ_E_NCCL;
    }
    *out = c;
    return 0;
}

// Enqueue all-reduce(sum) of buf[0..count) (device fp64, in place) after everything already enqueued on
// `stream`; returns at once.  The result may be read by work enqueued on `stream` after msq_comm_join.
// The buffer must stay allocated and untouched until then (the collective runs on the communicator's side stream).
static int comm_allreduce(msq_comm* c, void* buf, int count, int nccl_type, msq_stream_t stream) {
    if (!c || !buf || count < 1) return MSQ_E_BADARG;
    if (((uintptr_t)buf) & 7u) return MSQ_E_ALIGN;
    cudaError_t e;
    if ((e = cudaEventRecord(c->fork, (cudaStream_t)stream)) != cudaSuccess) return (int)e;
    if ((e = cudaStreamWaitEvent(c->side, c->fork, 0)) != cudaSuccess) return (int)e;
    if (nccl().all_reduce(buf, buf, (size_t)count, nccl_type, kNcclSum, c->comm, c->side) != 0) return MSQ_E_NCCL;
    if ((e = cudaEventRecord(c->done[c->issued % kRing], c->side)) != cudaSuccess) return (int)e;
    c->issued++;
    return 0;
}

extern "C" int msq_comm_allreduce_f64(msq_comm* c, double* buf, int count, msq_stream_t stream) {
    return comm_allreduce(c, buf, count, kNcclFloat64, stream);
}

// The integer results of the path -- the C x C confusion counts of Eval (utils/eval.py:121), the {cross-entropy sum in
// 2^-32 fixed point, valid-pixel count} pair of the guidance / source cross-entropy rows -- all-reduced as the uint64
// words they are: exact, and independent of the order NCCL adds them in.  Same stream semantics as the fp64 call.
extern "C" int msq_comm_allreduce_u64(msq_comm* c, unsigned long long* buf, int count, msq_stream_t stream) {
    return comm_allreduce(c, buf, count, kNcclUint64, stream);
}

// Make `stream` wait for the all-reduce issued `lag` calls before the most recent one (lag 0 = the most
// recent; no host synchronisation).  A lag of 1-2 lets a collective take more than one step without stalling
// the kernels: the statistics it carries are only logged.  On the mailbox path lag 0 also completes the steps in flight.
// Returns MSQ_E_PEER (once per loss) if a peer's vector did not arrive within the time-out since the last call.
extern "C" int msq_comm_join(msq_comm* c, int lag, msq_stream_t stream) {
    if (!c || lag < 0 || lag >= kRing) return MSQ_E_BADARG;
    if (c->box_ready && lag == 0 && c->reduced < c->seq) {
        // peer-memory path: the newest step's vector is not pushed yet, the one before it not reduced yet
        const unsigned K = c->seq;
        msq::PeerBox b = {};
        b.st = c->box_static;
        if (c->pushed < K) { b.cur = vec_slot(c, K); b.seq = K; b.count = (short)c->count[K % msq::kBoxSlots]; }
        if (K >= 2 && c->reduced < K - 1) { b.prev_out = red_slot(c, K - 1); b.prev_seq = K - 1; b.prev_count = (short)c->count[(K - 1) % msq::kBoxSlots]; }
        box_flush_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(b, K, c->count[K % msq::kBoxSlots], red_slot(c, K));
        c->pushed = K;
        c->reduced = K;
        MSQ_CHECK_LAUNCH();
        msq::count_launch();
    }
    if (c->issued > (unsigned long long)lag) {
        const cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, c->done[(c->issued - 1 - lag) % kRing], 0);
        if (e != cudaSuccess) return (int)e;
    }
    return peer_error(c);
}

// The all-reduced statistics vector of the msq_fused_fwd_bwd step issued `lag` steps before the most recent one
// (0 = the most recent), copied device-to-device into dst on `stream`.  It exists once two further steps have been
// enqueued (mailboxes), or after msq_comm_join(comm, 0, stream); MSQ_E_NOTREADY otherwise.
extern "C" int msq_comm_result(msq_comm* c, int lag, double* dst, int count, msq_stream_t stream) {
    if (!c || !dst || lag < 0 || lag >= msq::kBoxSlots - 1 || count < 1 || count > msq::kBoxCount) return MSQ_E_BADARG;
    if (c->seq <= (unsigned)lag) return MSQ_E_NOTREADY;
    const unsigned k = c->seq - (unsigned)lag;
    if (count > c->count[k % msq::kBoxSlots]) return MSQ_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (c->box_ready) {
        if (c->reduced < k) return MSQ_E_NOTREADY;
    } else if (c->world > 1) {
        const unsigned long long idx = c->ar_index[k % msq::kBoxSlots];
        if (c->issued - idx > (unsigned long long)kRing) return MSQ_E_NOTREADY;          // its event was recycled
        const cudaError_t e = cudaStreamWaitEvent(s, c->done[idx % kRing], 0);
        if (e != cudaSuccess) return (int)e;
    }
    return (int)cudaMemcpyAsync(dst, red_slot(c, k), (size_t)count * sizeof(double), cudaMemcpyDeviceToDevice, s);
}

// Sum over the ranks of `count` (<= 2) uint64 words, available in the SAME step -- the adjacent [ce_fix_out | nvalid_out] pair of
// the cross-entropy rows, which the head-2 / source backward divides by (msq_guidance_bwd): begin right after the forward,
// put independent work (the head-1 backward) on the stream, end right before the consumer.  With the mailboxes open: two
// 32-thread kernels, the words cross NVLink as 16-byte stores and are summed as integers by a spinning warp (a few us,
// against ~25 us for launching a collective); otherwise: one ncclAllReduce(uint64) on the side stream, joined by _end.
// One exchange may be in flight per communicator.  dst may equal src.
extern "C" int msq_comm_sum_u64_begin(msq_comm* c, const unsigned long long* src, int count, msq_stream_t stream) {
    if (!c || !src || count < 1 || count > kU64Max || c->u64_pending) return MSQ_E_BADARG;
    if (((uintptr_t)src) & 7u) return MSQ_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    c->u64_seq++;
    c->u64_pending = count;
    if (c->world == 1) return 0;
    if (c->box_ready) {
        u64_push_kernel<<<1, 32, 0, s>>>(c->box_static, c->u64_seq, src, count);
        MSQ_CHECK_LAUNCH();
        msq::count_launch();
        return 0;
    }
    unsigned long long* stage = c->u64_buf + (size_t)(c->u64_seq % kU64Ring) * kU64Max;
    const cudaError_t e = cudaMemcpyAsync(stage, src, sizeof(unsigned long long) * count, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return (int)e;
    return comm_allreduce(c, stage, count, kNcclUint64, stream);
}

extern "C" int msq_comm_sum_u64_end(msq_comm* c, unsigned long long* dst, int count, msq_stream_t stream) {
    if (!c || !dst || count != c->u64_pending || count < 1) return MSQ_E_BADARG;
    if (((uintptr_t)dst) & 7u) return MSQ_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    c->u64_pending = 0;
    if (c->world == 1) return 0;                     // the caller's words are the sum
    if (c->box_ready) {
        u64_reduce_kernel<<<1, 32, 0, s>>>(c->box_static, c->u64_seq, dst, count);
        MSQ_CHECK_LAUNCH();
        msq::count_launch();
        return peer_error(c);
    }
    const cudaError_t e = cudaStreamWaitEvent(s, c->done[(c->issued - 1) % kRing], 0);
    if (e != cudaSuccess) return (int)e;
    unsigned long long* stage = c->u64_buf + (size_t)(c->u64_seq % kU64Ring) * kU64Max;
    return (int)cudaMemcpyAsync(dst, stage, sizeof(unsigned long long) * count, cudaMemcpyDeviceToDevice, s);
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
        constexpr size_t kTail = sizeof(msq::PeerBoxStatic);
        if ((e = cudaMalloc((void**)&c->box_local, kBoxBytes + kTail)) != cudaSuccess) return (int)e;
        if ((e = cudaMemset(c->box_local, 0, kBoxBytes + kTail)) != cudaSuccess) return (int)e;
        if ((e = cudaHostAlloc((void**)&c->err_host, 64, cudaHostAllocMapped)) != cudaSuccess) return (int)e;
        memset(c->err_host, 0, 64);
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) return (int)e;
        c->box_static = (msq::PeerBoxStatic*)((char*)c->box_local + kBoxBytes);
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
    msq::PeerBoxStatic& st = c->box_static_host;
    st = msq::PeerBoxStatic{};
    st.world = c->world;
    st.rank = c->rank;
    cudaError_t e = cudaHostGetDevicePointer((void**)&st.err, c->err_host, 0);
    if (e != cudaSuccess) return (int)e;
    st.timeout_ns = default_timeout_ns();
    for (int p = 0; p < c->world; ++p) st.peer[p] = c->box_peer[p];
    if ((e = cudaMemcpy(c->box_static, &st, sizeof(st), cudaMemcpyHostToDevice)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return (int)e;
    c->box_mapped = true;
    return 0;
}

// How long one reduction waits for a peer's vector before giving that vector up (NaN statistics for that step, MSQ_E_PEER
// from the next msq_comm_join / msq_fused_fwd_bwd).  Default 600 s (MSQ_BOX_TIMEOUT_S), the order of NCCL's watchdog: rank
// skew of many seconds is routine (rank-0-only validation, checkpoints, a dataloader stall).  Synchronises the device.
extern "C" int msq_comm_box_timeout(msq_comm* c, double seconds) {
    if (!c || !(seconds > 0.0) || !c->box_mapped) return MSQ_E_BADARG;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    c->box_static_host.timeout_ns = (unsigned long long)(seconds * 1e9);
    e = cudaMemcpy(c->box_static, &c->box_static_host, sizeof(msq::PeerBoxStatic), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaDeviceSynchronize();
}

// Switch the mailbox path on (only if every peer is mapped) or off.  The caller enables it on all ranks or on none:
// after msq_comm_box_open it agrees on the outcome across the ranks (an all-reduce(min) of "my open succeeded").
extern "C" int msq_comm_box_enable(msq_comm* c, int on) {
    if (!c) return MSQ_E_BADARG;
    if (on && !c->box_mapped) return MSQ_E_BADARG;
    if (c->box_ready && c->reduced < c->seq) return MSQ_E_BADARG;          // steps in flight: msq_comm_join first
    c->box_ready = on != 0;
    c->pushed = c->reduced = c->seq;
    return 0;
}

// 1 if msq_fused_fwd_bwd exchanges the statistics through the mailboxes, 0 if it uses ncclAllReduce
extern "C" int msq_comm_box_active(const msq_comm* c) { return (c && c->box_ready) ? 1 : 0; }

// error words of the mailbox path: *out bit 0 = a peer's vector did not arrive in time at least once; read from mapped
// host memory after synchronising the device
extern "C" int msq_comm_box_errors(msq_comm* c, unsigned* out) {
    if (!c || !out) return MSQ_E_BADARG;
    *out = 0u;
    if (!c->err_host) return 0;
    const cudaError_t e = cudaDeviceSynchronize();
    *out = ((volatile unsigned*)c->err_host)[0];
    return (int)e;
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
    if (c->err_host) cudaFreeHost(c->err_host);
    if (c->ring) cudaFree(c->ring);
    if (c->u64_buf) cudaFree(c->u64_buf);
    if (c->comm) nccl().comm_destroy(c->comm);
    cudaEventDestroy(c->fork);
    for (int i = 0; i < kRing; ++i) cudaEventDestroy(c->done[i]);
    cudaStreamDestroy(c->side);
    delete c;
}

// One call per training step for callers that know the upstream gradient scale when they call the forward
// (lambda_target is a constant, tools/solve_gta5.py:199,217): the results of msq_fused_fwd + msq_fused_bwd from TWO kernels
// (the backward derives the weights itself and carries the finalisation in an extra CTA, fused_common.cuh) and, when the
// images are sharded over ranks (comm != NULL), the exchange of the step's statistics vector: carried by a second extra
// CTA of the backward over the peer-memory mailboxes when they are open, else one ncclAllReduce forked after the backward
// (so that nothing sits between the step's kernels) and ordered after the collective issued `lag` steps earlier.  It also
// keeps the host side of a 29 us step to one library call.
// `out.stats` keeps this rank's LOCAL vector; the all-reduced one is kept by the communicator (msq_comm_result).
// All steps of one communicator must be enqueued on the same stream (the mailbox protocol relies on stream order).
extern "C" int msq_fused_fwd_bwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                                 double ratio, int n_images_norm, void* accum, void* out, void* aux, const float* grad_out,
                                 float grad_scale, float* grad_logits, msq_comm* comm, int lag, msq_stream_t stream) {
    if (!grad_logits || !out || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES) return MSQ_E_BADARG;
    if ((((uintptr_t)aux) & 15u) || (((uintptr_t)grad_logits) & 3u)) return MSQ_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    const int nstat = 1 + num_class;
    if (comm && comm->box_ready && nstat <= msq::kBoxCount) {
        // sharded: an extra CTA of this step's backward kernel (of its finalisation kernel when late_finalize = 0) pushes the PREVIOUS step's [loss | hist] into every
        // rank's mailbox over NVLink and reduces the step before that; nothing is enqueued between or after the three
        // kernels of the step and the hot kernels are untouched
        const unsigned k = comm->seq + 1u;
        msq::PeerBox b = {};
        b.st = comm->box_static;
        b.keep = vec_slot(comm, k);
        b.keep_count = (short)nstat;
        if (k >= 2 && comm->pushed < k - 1) { b.cur = vec_slot(comm, k - 1); b.seq = k - 1; b.count = (short)comm->count[(k - 1) % msq::kBoxSlots]; }
        if (k >= 3 && comm->reduced < k - 2) { b.prev_out = red_slot(comm, k - 2); b.prev_seq = k - 2; b.prev_count = (short)comm->count[(k - 2) % msq::kBoxSlots]; }
        const int late = msq::g_late_finalize;          // forward -> backward (derives the weights; finalisation + exchange in extra CTAs)
        int rc = msq::fused_fwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, nullptr, ratio, n_images_norm, accum,
                                         out, aux, grad_logits, s, 0, &b, late);
        if (rc) return rc;
        comm->seq = k;
        comm->count[k % msq::kBoxSlots] = nstat;
        if (b.cur) comm->pushed = k - 1;
        if (b.prev_out) comm->reduced = k - 2;
        rc = msq::fused_bwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, n_images_norm, out, grad_out, grad_scale,
                                     grad_logits, aux, 1, s, 0, late ? accum : nullptr, ratio, &b);
        return rc ? rc : peer_error(comm);
    }
    const int late = msq::g_late_finalize;
    int rc = msq::fused_fwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, nullptr, ratio, n_images_norm, accum, out,
                                     aux, grad_logits, s, 0, nullptr, late);
    if (rc) return rc;
    rc = msq::fused_bwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, n_images_norm, out, grad_out, grad_scale,
                                 grad_logits, aux, 1, s, 0, late ? accum : nullptr, ratio, nullptr);
    if (rc || !comm || nstat > msq::kBoxCount) return rc;
    // NCCL path: the vector is copied into the communicator's ring on the caller's stream (after the backward: nothing
    // between the step's kernels) and all-reduced THERE on the side stream; the caller's `out` is never touched later
    const msq_state_layout lay = msq::make_layout(n, num_class);
    const unsigned k = comm->seq + 1u;
    cudaError_t e = cudaMemcpyAsync(red_slot(comm, k), (const char*)out + lay.stats_off, (size_t)nstat * sizeof(double),
                                    cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return (int)e;
    comm->seq = k;
    comm->count[k % msq::kBoxSlots] = nstat;
    if (comm->world == 1) return 0;
    rc = msq_comm_join(comm, lag, stream);
    if (rc) return rc;
    comm->ar_index[k % msq::kBoxSlots] = comm->issued;
    return msq_comm_allreduce_f64(comm, red_slot(comm, k), nstat, stream);
}
