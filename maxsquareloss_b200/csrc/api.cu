// ABI housekeeping: version, error names, state layout, tuning knobs.
#include <string.h>
#include "common.cuh"

namespace msq {
extern int g_conf_agg;      // confusion.cu
extern int g_conf_ctas_per_sm;
extern int g_prob_waves;    // prob_loss.cu
extern int g_fused_rows;    // fused_loss.cu
}

namespace msq {

// one warp per image, one lane per class (C <= 32)
__global__ void __launch_bounds__(256)
finalize_kernel(State st, int mode, int n, int C, float r32, float omr32, int n_norm, unsigned long long kept_dense,
                int multi) {
    __shared__ double s_red[8];
    __shared__ unsigned long long s_cls[MSQ_MAX_CLASSES];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
    pdl_trigger();          // the backward may start its prologue now
    pdl_wait();             // ... but this kernel needs every forward CTA's atomics
    // thread 0's scalars are fetched first so that their latency overlaps everything else
    unsigned long long kept_ld = 0ull;
    unsigned flags_ld = 0u;
    if (tid == 0) { kept_ld = *st.kept; flags_ld = *st.flags; }
    // multi-level guidance: cross-entropy sum and valid-pixel count, one replica per lane of warp 0
    unsigned long long ce_ld = 0ull, nv_ld = 0ull;
    if (multi && tid < kRep) { ce_ld = st.ce[tid]; nv_ld = st.nvalid[tid]; st.ce[tid] = 0ull; st.nvalid[tid] = 0ull; }
    if (tid < MSQ_MAX_CLASSES) s_cls[tid] = 0ull;
    __syncthreads();
    const int nc = n * C;
    double part = 0.0;
    unsigned long long cls_tot = 0ull;
    for (int img = wid; img < n; img += nw) {
        const int idx = img * C + lane;
        unsigned hcnt = 0u;
        unsigned long long sq = 0ull;
        if (lane < C) {
#pragma unroll
            for (int r = 0; r < kRep; ++r) {          // independent loads: one latency
                hcnt += st.hist[r * nc + idx];
                sq += st.sumsq[r * nc + idx];
            }
#pragma unroll
            for (int r = 0; r < kRep; ++r) { st.hist[r * nc + idx] = 0u; st.sumsq[r * nc + idx] = 0ull; }   // self-clean
        }
        const unsigned total = __reduce_add_sync(0xffffffffu, hcnt);
        const double S = (double)sq * kInvFix;
        float wgt = 1.0f;
        if (mode == MSQ_MODE_IW && lane < C) wgt = iw_weight((float)hcnt, (float)total, r32, omr32);
        if (lane < C) {
            part += (double)wgt * S;
            st.weights[idx] = wgt;
            st.hist_out[idx] = (int)hcnt;
            cls_tot += hcnt;
        }
        double simg = S;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) simg += __shfl_xor_sync(0xffffffffu, simg, o);
        if (lane == 0) st.sum_out[img] = simg;
    }
    if (multi && wid == 0) {
        ce_ld = warp_sum_u64(ce_ld);
        nv_ld = warp_sum_u64(nv_ld);
    }
    if (lane < C && cls_tot) atomicAdd(&s_cls[lane], cls_tot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[wid] = part;
    __syncthreads();
    if (tid < C) st.stats[1 + tid] = (double)s_cls[tid];
    if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < nw; ++i) tot += s_red[i];
        const unsigned long long kept_local = (kept_dense != 0ull) ? kept_dense : kept_ld;
        const double scale = (double)n_norm / (double)n;
        double loss;
        if (mode == MSQ_MODE_IW) loss = -tot / ((double)n_norm * (double)C);
        else loss = -tot / (2.0 * (double)kept_local * scale);
        if (flags_ld & kFlagNonFinite) loss = __longlong_as_double(0x7ff8000000000000LL);
        *st.loss = (float)loss;
        *st.kept_out = kept_local;
        st.stats[0] = loss;
        if (multi) {       // nn.CrossEntropyLoss(ignore_index=-1): mean over the valid pixels, 0/0 = NaN as in torch
            double l2 = ((double)ce_ld * kInvFix) / (double)nv_ld;
            if (flags_ld & kFlagNonFinite) l2 = __longlong_as_double(0x7ff8000000000000LL);
            *st.loss2 = (float)l2;
            *st.nvalid_out = nv_ld;
            *st.ce_out = (double)ce_ld * kInvFix;
        }
        *st.kept = 0ull;                       // self-clean
        *st.flags = 0u;
    }
}

int launch_finalize(const State& st, int mode, int n, int C, float r32, float omr32, int n_norm,
                    unsigned long long kept_dense, cudaStream_t stream, int multi) {
    const int warps = n < 8 ? n : 8;
    const cudaError_t e = launch_pdl(finalize_kernel, dim3(1), dim3(32 * warps), 0, stream, st, mode, n, C, r32, omr32,
                                     n_norm, kept_dense, multi);
    if (e != cudaSuccess) return (int)e;
    MSQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace msq

extern "C" int msq_abi_version(void) { return MSQ_ABI_VERSION; }

extern "C" const char* msq_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case MSQ_E_BADARG: return "msq: bad argument (null pointer, non-positive size or more than 32 classes)";
        case MSQ_E_GEOMETRY: return "msq: fused path needs out_h >= h and out_w >= w (bilinear upsampling only)";
        case MSQ_E_SMEM: return "msq: low-resolution tile does not fit in shared memory";
        case MSQ_E_ALIGN: return "msq: pointer is not aligned for its element type";
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
//   "prob_waves" W      grid of the strict kernels = W x co-resident capacity
//   "fused_rows" R      about R output rows per CTA in the fused kernels (0 = automatic: one balanced wave)
extern "C" int msq_tune_set(const char* key, int value) {
    if (!key) return MSQ_E_BADARG;
    if (!strcmp(key, "conf_agg")) { msq::g_conf_agg = value; return 0; }
    if (!strcmp(key, "conf_ctas")) { msq::g_conf_ctas_per_sm = value; return 0; }
    if (!strcmp(key, "prob_waves")) { msq::g_prob_waves = value; return 0; }
    if (!strcmp(key, "fused_rows")) { msq::g_fused_rows = value; return 0; }
    return MSQ_E_BADARG;
}
