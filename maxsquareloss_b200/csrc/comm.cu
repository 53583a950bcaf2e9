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
};

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
    if (c->issued <= (unsigned long long)lag) return 0;
    const cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, c->done[(c->issued - 1 - lag) % kRing], 0);
    return (int)e;
}

extern "C" void msq_comm_destroy(msq_comm* c) {
    if (!c) return;
    cudaStreamSynchronize(c->side);
    if (c->comm) nccl().comm_destroy(c->comm);
    cudaEventDestroy(c->fork);
    for (int i = 0; i < kRing; ++i) cudaEventDestroy(c->done[i]);
    cudaStreamDestroy(c->side);
    delete c;
}

// One call per training step for callers that know the upstream gradient scale when they call the forward
// (lambda_target is a constant, tools/solve_gta5.py:199,217): msq_fused_fwd + msq_fused_bwd and, when the images are
// sharded over ranks (comm != NULL), the step's statistics all-reduce -- forked after the backward so that nothing
// sits between forward -> finalise -> backward, and ordered after the collective issued `lag` steps earlier.  Same
// kernels and results as the separate calls; it exists to keep the host side of a 35 us step to one library call.
extern "C" int msq_fused_fwd_bwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                                 double ratio, int n_images_norm, void* accum, void* out, void* aux, const float* grad_out,
                                 float grad_scale, float* grad_logits, msq_comm* comm, int lag, msq_stream_t stream) {
    if (!grad_logits) return MSQ_E_BADARG;
    if ((((uintptr_t)aux) & 15u) || (((uintptr_t)grad_logits) & 3u)) return MSQ_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = msq::fused_fwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, nullptr, ratio, n_images_norm, accum, out,
                                     aux, grad_logits, s);
    if (rc) return rc;
    rc = msq::fused_bwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, n_images_norm, out, grad_out, grad_scale,
                                 grad_logits, aux, 1, s);
    if (rc || !comm) return rc;
    rc = msq_comm_join(comm, lag, stream);
    if (rc) return rc;
    const msq_state_layout lay = msq::make_layout(n, num_class);
    return msq_comm_allreduce_f64(comm, (double*)((char*)out + lay.stats_off), 1 + num_class, stream);
}
