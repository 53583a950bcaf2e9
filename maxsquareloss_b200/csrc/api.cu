// ABI housekeeping: version, error names, state layout, tuning knobs.
#include <string.h>
#define MSQ_TRACE_FINALIZE_BODY 1
#include "common.cuh"

namespace msq {
extern int g_conf_agg;      // confusion.cu
extern int g_conf_ctas_per_sm;
extern int g_conf_grid;
extern int g_prob_waves;    // prob_loss.cu
extern int g_fused_rows;    // fused_loss.cu
extern int g_reserve_sms;
}

namespace msq {

// one warp per image, one lane per class (C <= 32)
__global__ void __launch_bounds__(256)
finalize_kernel(State st, int mode, int n, int C, float r32, float omr32, int n_norm, unsigned long long kept_dense,
                int multi, int loss_kind, const PeerBox box) {
    MSQ_TRACE_PT(0, 0);
    pdl_trigger();          // the backward may start its prologue now
    pdl_wait();             // ... but this kernel needs every forward CTA's atomics
    MSQ_TRACE_PT(0, 1);
    if (blockIdx.x == 1) {  // sharded step: the statistics exchange over NVLink peer memory rides along (PeerBox, common.cuh)
        if (threadIdx.x < 32) box_exchange(box, (int)threadIdx.x);
        return;
    }
    finalize_body(st, mode, n, C, r32, omr32, n_norm, kept_dense, multi, loss_kind);
    if (box.keep) {         // this step's own [loss | hist] into the communicator's ring: pushed by the NEXT step's exchange
        __syncthreads();
        for (int k = threadIdx.x; k < box.keep_count; k += blockDim.x) box.keep[k] = st.stats[k];
    }
    MSQ_TRACE_PT(0, 2);
}

int launch_finalize(const State& st, int mode, int n, int C, float r32, float omr32, int n_norm,
                    unsigned long long kept_dense, cudaStream_t stream, int multi, int loss_kind, const PeerBox* box) {
    // at most 4 warps, the warp count of the backward CTA that runs the same body in the one-call step (fin_cta): the fp64
    // partial sums are then added in the same order on both paths, whatever the number of images (bit-identical losses)
    const int warps = n < 4 ? n : 4;
    const PeerBox bx = box ? *box : PeerBox{};
    const cudaError_t e = launch_pdl_as(2, finalize_kernel, dim3((bx.st && (bx.cur || bx.prev_out)) ? 2 : 1), dim3(32 * warps), 0, stream, st, mode, n, C, r32, omr32,
                                     n_norm, kept_dense, multi, loss_kind, bx);
    if (e != cudaSuccess) return (int)e;
    MSQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace msq

namespace msq { unsigned long long g_launches = 0ull; int g_pdl_mask = 15; }

#if MSQ_TRACE
extern "C" int msq_debug_trace_finalize(unsigned long long* host_dst, long long count) {
    return (int)cudaMemcpyFromSymbol(host_dst, msq::g_trace, (size_t)count * 8);
}
#endif

extern "C" int msq_abi_version(void) { return MSQ_ABI_VERSION; }

extern "C" unsigned long long msq_launch_count(void) { return __atomic_load_n(&msq::g_launches, __ATOMIC_RELAXED); }

extern "C" const char* msq_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case MSQ_E_BADARG: return "msq: bad argument (null pointer, non-positive size or more than 32 classes)";
        case MSQ_E_GEOMETRY: return "msq: fused path needs out_h >= h and out_w >= w (bilinear upsampling only)";
        case MSQ_E_SMEM: return "msq: low-resolution tile does not fit in shared memory";
        case MSQ_E_ALIGN: return "msq: pointer is not aligned for its element type";
        case MSQ_E_NCCL: return "msq: libnccl.so.2 could not be loaded, or an NCCL call failed";
        case MSQ_E_PEER: return "msq: a peer's statistics vector did not arrive within the mailbox time-out (that step's all-reduced statistics are NaN)";
        case MSQ_E_NOTREADY: return "msq: that step's all-reduced statistics do not exist yet (two steps later, or after msq_comm_join)";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "msq: unknown error";
    }
}

extern "C" int msq_state_layout_get(int n_images, int num_class, msq_state_layout* out) {
    if (!out || n_images < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES) return MSQ_E_BADARG;
    *out = msq::make_layout(n_images, num_class);
    return 0;
}

// Performance-tuning knobs (bench sweeps); results never depend on them.
//   "conf_agg"   0|1|2  warp aggregation level of the confusion histogram
//   "conf_ctas"  1|2    1024-thread CTAs per SM of the int64 confusion kernel
//   "conf_grid"  G      at most G CTAs in the int64 confusion kernel (0 = automatic)
//   "prob_waves" W      grid of the strict kernels = W x co-resident capacity
//   "fused_rows" R      about R output rows per CTA in the fused kernels (0 = automatic: one balanced wave)
//   "reserve_sms" S     the one-wave grids of the fused kernels leave S SMs free (a concurrent NCCL kernel gets them)
//   "late_finalize" 0|1 msq_fused_fwd_bwd: two kernels, forward -> backward that derives the weights itself and carries the
//                       finalisation in an extra CTA (default 1), or forward -> finalisation -> backward (0)
//   "pdl_mask"   M      which launches carry the programmatic-stream-serialisation attribute: bit 0 fused forward, bit 1
//                       finalisation, bit 2 fused backward, bit 3 every other kernel (default 15 = all)
extern "C" int msq_tune_set(const char* key, int value) {
    if (!key) return MSQ_E_BADARG;
    if (!strcmp(key, "conf_agg")) { msq::g_conf_agg = value; return 0; }
    if (!strcmp(key, "conf_ctas")) { msq::g_conf_ctas_per_sm = value; return 0; }
    if (!strcmp(key, "conf_grid")) { msq::g_conf_grid = value; return 0; }
    if (!strcmp(key, "prob_waves")) { msq::g_prob_waves = value; return 0; }
    if (!strcmp(key, "fused_rows")) { msq::g_fused_rows = value; return 0; }
    if (!strcmp(key, "reserve_sms")) { msq::g_reserve_sms = value; return 0; }
    if (!strcmp(key, "pdl_mask")) { msq::g_pdl_mask = value; return 0; }
    if (!strcmp(key, "late_finalize")) { msq::g_late_finalize = value; return 0; }
    return MSQ_E_BADARG;
}
