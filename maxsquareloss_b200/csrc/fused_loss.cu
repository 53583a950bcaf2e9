// Fused kernels: low-resolution head logits in, loss / histogram / dL/dlogits out.
//
// Absorbs, per output pixel and without materialising anything at label
// resolution:
//   F.interpolate(x, size, mode='bilinear', align_corners=True)   graphs/models/deeplab_multi.py:124,128
//   F.softmax(pred, dim=1)                                        tools/solve_gta5.py:182-183
//   MaxSquareloss / IW_MaxSquareloss forward                      utils/loss.py:76-102,110-119
// and, in the backward kernel, the adjoint of all three.
//
// Work decomposition ("column walk"): a CTA owns a TW-column x R-row strip of
// the OUTPUT image; one thread owns one output column and walks down the rows.
// ATen's bilinear formula is horizontal-first,
//     t_r = fma(A[r,x0], lx0, A[r,x1]*lx1);   z = fma(t_y0, ly0, t_y1*ly1)
// (pinned bit-for-bit in oracle/bilinear.py), and for a fixed column the
// horizontally interpolated values t_r of a low-res row r are shared by all
// ~H/h output rows that use r -- so the thread keeps t_y0[c], t_y1[c] (2C
// registers) and refreshes one of them only when y0 advances.  Per pixel that
// leaves 2 FP ops per class for the upsample and NO memory traffic: the
// low-res tile the strip needs is staged once in shared memory.  The backward
// kernel mirrors this with two register accumulators d t_y0[c], d t_y1[c] per
// thread that are flushed through a shared-memory transpose (horizontal
// adjoint, one value per low-res cell) and one global red.add per cell when y0
// advances.  These kernels are FP32-issue / MUFU bound, not HBM bound (3.65
// algorithmic bytes per pixel at C=19, 65x129 -> 512x1024).
#include "fused_common.cuh"

namespace msq {

#if MSQ_TRACE
static unsigned g_trace_par = 0u;       // step parity, bit 31 of the kernels' `units` argument (scripts/trace_step.py)
#define MSQ_UNITS(u) ((unsigned)(u) | (g_trace_par << 31))
#else
#define MSQ_UNITS(u) ((unsigned)(u))
#endif
#ifndef MSQ_BWD_SPARE
#define MSQ_BWD_SPARE 1                         // CTA slots the backward leaves to the finalisation kernel (fused_common.cuh, plan_launch)
#endif
int g_late_finalize = 1;
int g_fused_rows = 0;
int g_reserve_sms = 0;

template <int CT, bool PAD>
static int launch_fused_fwd(int mode, const float* lo, int C, int h, int w, int H, int W, int n, const int64_t* label,
                            float r32, float omr32, int nn, State st, void* aux, float* zero_buf, cudaStream_t s, int loss_kind,
                            const PeerBox* box, int late_finalize) {
    const unsigned zero_count = zero_buf ? (unsigned)((size_t)n * C * h * w) : 0u;
    const bool iw = mode != MSQ_MODE_MAXSQUARE;
#if MSQ_TRACE
    g_trace_par ^= 1u;
#endif
#define MSQ_LAUNCH(K)                                                                          \
    do {                                                                                       \
        LaunchPlan lp;                                                                         \
        const int rc = plan_launch(K, C, h, w, H, W, n, MSQ_FWD_MINB,                          \
                                   [&](const FusedGeo& g) { return fwd_smem(g, iw, CT); }, lp); \
        if (rc) return rc;                                                                     \
        const cudaError_t le = launch_pdl_as(1, K, dim3(lp.p.grid), dim3(kTW), lp.smem, s, lo, lp.p.g, n, MSQ_UNITS(lp.p.units), \
                                          label, st, aux, zero_buf, zero_count);              \
        if (le != cudaSuccess) return (int)le;                                                 \
    } while (0)
    if (loss_kind == 1) {            // MinEnt losses: no label= argument in the reference (utils/loss.py:45)
        if (label) return MSQ_E_BADARG;
        if (!iw) MSQ_LAUNCH((fused_fwd_kernel<CT, PAD, false, false, 1>));
        else MSQ_LAUNCH((fused_fwd_kernel<CT, PAD, true, false, 1>));
    } else if (!iw) MSQ_LAUNCH((fused_fwd_kernel<CT, PAD, false, false>));
    else if (label) MSQ_LAUNCH((fused_fwd_kernel<CT, PAD, true, true>));
    else MSQ_LAUNCH((fused_fwd_kernel<CT, PAD, true, false>));
#undef MSQ_LAUNCH
    MSQ_CHECK_LAUNCH();
    if (late_finalize) return 0;     // the one-call step runs it in extra CTAs of the backward (fin_cta, fused_common.cuh)
    return launch_finalize(st, mode, n, C, r32, omr32, nn, (unsigned long long)n * C * H * W, s, 0, loss_kind, box);
}

template <int CT, bool PAD>
static int launch_fused_bwd(int mode, const float* lo, int C, int h, int w, int H, int W, int n, int nn, State st,
                            const float* grad_out, float grad_out_value, float* grad_lo, const void* aux,
                            bool grad_is_zeroed, cudaStream_t s, int loss_kind, const FinArgs& fin) {
    if (!grad_is_zeroed) {
        cudaError_t e = cudaMemsetAsync(grad_lo, 0, (size_t)n * C * h * w * sizeof(float), s);
        if (e != cudaSuccess) return (int)e;
    }
#define MSQ_LAUNCH(K)                                                                          \
    do {                                                                                       \
        LaunchPlan lp;                                                                         \
        const int rc = plan_launch(K, C, h, w, H, W, n, MSQ_BWD_MINB,                          \
                                   [&](const FusedGeo& g) { return bwd_smem(g, CT); }, lp, fin.extra > MSQ_BWD_SPARE ? fin.extra : MSQ_BWD_SPARE);    \
        if (rc) return rc;                                                                     \
        const cudaError_t le = launch_pdl_as(4, K, dim3(lp.p.grid + fin.extra), dim3(kTW), lp.smem, s, lo, lp.p.g, n, MSQ_UNITS(lp.p.units), nn, \
                                          (const float*)st.weights, grad_out, grad_out_value, grad_lo, aux,           \
                                          (const unsigned long long*)nullptr, fin);                                  \
        if (le != cudaSuccess) return (int)le;                                                                        \
    } while (0)
    if (loss_kind == 1) {            // MinEnt: the backward always replays the forward's cache
        if (!aux) return MSQ_E_BADARG;
        if (mode == MSQ_MODE_MAXSQUARE) MSQ_LAUNCH((fused_bwd_kernel<CT, PAD, false, true, false, 1>));
        else MSQ_LAUNCH((fused_bwd_kernel<CT, PAD, true, true, false, 1>));
    } else if (mode == MSQ_MODE_MAXSQUARE) {
        if (aux) MSQ_LAUNCH((fused_bwd_kernel<CT, PAD, false, true>));
        else MSQ_LAUNCH((fused_bwd_kernel<CT, PAD, false, false>));
    } else {
        if (aux) MSQ_LAUNCH((fused_bwd_kernel<CT, PAD, true, true>));
        else MSQ_LAUNCH((fused_bwd_kernel<CT, PAD, true, false>));
    }
#undef MSQ_LAUNCH
    MSQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace msq

#define MSQ_DISPATCH_C(C, CALL)                  \
    switch (C) {                                 \
        case 13: return CALL(13, false);         \
        case 16: return CALL(16, false);         \
        case 19: return CALL(19, false);         \
        default:                                 \
            if ((C) <= 8) return CALL(8, true);  \
            if ((C) <= 24) return CALL(24, true);\
            return CALL(32, true);               \
    }

namespace msq {

int fused_fwd_dispatch(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                       const int64_t* label, double ratio, int n_images_norm, void* accum, void* out, void* aux,
                       float* zero_grad, cudaStream_t s, int loss_kind, const PeerBox* box, int late_finalize) {
    if (!logits || !accum || !out || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES || h < 1 || w < 1 || out_h < 1 ||
        out_w < 1)
        return MSQ_E_BADARG;
    if (mode != MSQ_MODE_IW && mode != MSQ_MODE_MAXSQUARE) return MSQ_E_BADARG;
    if ((((uintptr_t)logits) & 3u) || ((((uintptr_t)accum) | ((uintptr_t)out)) & 15u) || (label && (((uintptr_t)label) & 7u))) return MSQ_E_ALIGN;
    const State st = carve(accum, out, n, num_class);
    const float r32 = (float)ratio, omr32 = (float)(1.0 - ratio);
    const int nn = n_images_norm > 0 ? n_images_norm : n;
#define CALL(CT, PAD) launch_fused_fwd<CT, PAD>(mode, logits, num_class, h, w, out_h, out_w, n, label, r32, omr32, nn, st, aux, zero_grad, s, loss_kind, box, late_finalize)
    MSQ_DISPATCH_C(num_class, CALL)
#undef CALL
}

int fused_bwd_dispatch(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                       int n_images_norm, const void* out, const float* grad_out, float grad_out_value,
                       float* grad_logits, const void* aux, int grad_is_zeroed, cudaStream_t s, int loss_kind,
                       void* accum_fin, double ratio, const PeerBox* box) {
    if (!logits || !out || !grad_logits || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES ||
        h < 1 || w < 1 || out_h < 1 || out_w < 1)
        return MSQ_E_BADARG;
    if (mode != MSQ_MODE_IW && mode != MSQ_MODE_MAXSQUARE) return MSQ_E_BADARG;
    if ((((uintptr_t)logits) | ((uintptr_t)grad_logits) | ((uintptr_t)grad_out)) & 3u) return MSQ_E_ALIGN;
    const State st = carve(accum_fin, const_cast<void*>(out), n, num_class);
    const int nn = n_images_norm > 0 ? n_images_norm : n;
    FinArgs fin = {};
    if (accum_fin) {                 // one-call step: the finalisation rides in extra CTAs of this launch
        fin.st = st;
        if (box) fin.box = *box;
        fin.kept_dense = (unsigned long long)n * num_class * out_h * out_w;
        fin.mode = mode; fin.n = n; fin.C = num_class; fin.n_norm = nn; fin.loss_kind = loss_kind;
        fin.r32 = (float)ratio; fin.omr32 = (float)(1.0 - ratio);
        fin.extra = (fin.box.st && (fin.box.cur || fin.box.prev_out)) ? 2 : 1;
    }
#define CALL(CT, PAD) launch_fused_bwd<CT, PAD>(mode, logits, num_class, h, w, out_h, out_w, n, nn, st, grad_out, grad_out_value, grad_logits, aux, grad_is_zeroed != 0, s, loss_kind, fin)
    MSQ_DISPATCH_C(num_class, CALL)
#undef CALL
}


}  // namespace msq

extern "C" int64_t msq_fused_aux_bytes(int n, int out_h, int out_w) {
    if (n < 1 || out_h < 1 || out_w < 1) return 0;
    const int64_t npix = (int64_t)n * out_h * out_w;
    return 16 * npix;
}

extern "C" int msq_fused_fwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                             const int64_t* label, double ratio, int n_images_norm, void* accum, void* out,
                             void* aux, float* zero_grad, msq_stream_t stream) {
    if ((((uintptr_t)aux) & 15u) || (((uintptr_t)zero_grad) & 3u)) return MSQ_E_ALIGN;
    return msq::fused_fwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, label, ratio, n_images_norm, accum, out,
                                   aux, zero_grad, (cudaStream_t)stream);
}

extern "C" int msq_fused_bwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                             int n_images_norm, const void* out, const void* aux, const float* grad_out,
                             float* grad_logits, int grad_is_zeroed, msq_stream_t stream) {
    if (!grad_out) return MSQ_E_BADARG;
    if (((uintptr_t)aux) & 15u) return MSQ_E_ALIGN;
    return msq::fused_bwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, n_images_norm, out, grad_out, 0.f,
                                   grad_logits, aux, grad_is_zeroed, (cudaStream_t)stream);
}

// MinEnt baselines of the same factory (tools/solve_gta5.py:150-155): softCrossEntropy (mode MAXSQUARE =
// unweighted) and IWsoftCrossEntropy (mode IW) of utils/loss.py:17-67, called as the trainers call them,
// with target = softmax(inputs) (tools/solve_gta5.py:188-190,199), fused from the low-resolution logits.
//   loss = mean(-p log p)                              (softCrossEntropy)
//   loss = sum_px w[argmax logits] * H_px / (N C)      (IWsoftCrossEntropy; H = -sum_c p_c log p_c)
// Same buffers and outputs as msq_fused_fwd / msq_fused_bwd; `aux` is required (the backward replays it).
extern "C" int msq_entropy_fwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                               double ratio, int n_images_norm, void* accum, void* out, void* aux, float* zero_grad,
                               msq_stream_t stream) {
    if ((((uintptr_t)aux) & 15u) || (((uintptr_t)zero_grad) & 3u)) return MSQ_E_ALIGN;
    return msq::fused_fwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, nullptr, ratio, n_images_norm, accum, out,
                                   aux, zero_grad, (cudaStream_t)stream, 1);
}

extern "C" int msq_entropy_bwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                               int n_images_norm, const void* out, const void* aux, const float* grad_out,
                               float* grad_logits, int grad_is_zeroed, msq_stream_t stream) {
    if (!grad_out || !aux) return MSQ_E_BADARG;
    if (((uintptr_t)aux) & 15u) return MSQ_E_ALIGN;
    return msq::fused_bwd_dispatch(mode, logits, n, num_class, h, w, out_h, out_w, n_images_norm, out, grad_out, 0.f,
                                   grad_logits, aux, grad_is_zeroed, (cudaStream_t)stream, 1);
}

#if MSQ_TRACE
extern "C" int msq_debug_trace_fused(unsigned long long* host_dst, long long count) {
    return (int)cudaMemcpyFromSymbol(host_dst, msq::g_trace, (size_t)count * 8);
}
#endif
