// Flip-ensemble evaluation, fused (tools/evaluate.py:120-141 with --flip):
//
//   pred_P   = softmax(model(x))                      pred_P_2 = flip(softmax(model(flip(x))), -1)
//   pred_c   = (pred_P + pred_P_2) / 2                argpred  = np.argmax(pred_c.cpu().numpy(), axis=1)
//   Eval.add_batch(label, argpred)                    utils/eval.py:109-121
//
// One pass over the two logits tensors (2 * 4C B/pixel + 8 B of label): both softmaxes, the mirrored
// average, the argmax and the confusion-matrix update; the reference moves 2 x 40 MB per image to
// the host and argmaxes there.  The argmax is an integer result and must match bit for bit: the
// fast path uses ex2.approx, and a pixel whose two best averaged probabilities are within 1e-5
// relative is re-evaluated with torch's own arithmetic (expf(z - m) / sum in class order,
// (p1 + p2) / 2 in fp32, first maximum wins).
#include <stdlib.h>
#include "common.cuh"

namespace msq {

constexpr int kFlipThreads = 256;

__device__ __forceinline__ int ensemble_exact(const float* za, const float* zb, int C, float ma, float mb) {
    float sa = 0.f, sb = 0.f;
    for (int c = 0; c < C; ++c) { sa += expf(za[c] - ma); sb += expf(zb[c] - mb); }
    float best = -1.f;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
        const float pa = __fdiv_rn(expf(za[c] - ma), sa), pb = __fdiv_rn(expf(zb[c] - mb), sb);
        const float pc = __fmul_rn(__fadd_rn(pa, pb), 0.5f);
        if (pc > best) { best = pc; arg = c; }
    }
    return arg;
}

// Fallback for odd widths / unaligned views and generic class counts: one thread per pixel; x runs along the image row so
// that both the direct read (x) and the mirrored read (W-1-x) of a warp are one contiguous 128-byte segment
template <int CT>
__global__ void __launch_bounds__(kFlipThreads)
confusion_flip_kernel(const int64_t* __restrict__ gt, const float* __restrict__ la, const float* __restrict__ lb, int C,
                      int H, int W, unsigned long long* __restrict__ cm) {
    extern __shared__ unsigned s_cm[];
    const int nbins = C * C;
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) s_cm[b] = 0u;
    __syncthreads();
    pdl_trigger();          // programmatic dependent launch (see confusion.cu)
    pdl_wait();
    const int n = blockIdx.y;
    const long long hw = (long long)H * W;
    const int64_t* gt_n = gt + (long long)n * hw;
    const float* a_n = la + (long long)n * C * hw;
    const float* b_n = lb + (long long)n * C * hw;
    constexpr int CR = CT > 0 ? CT : 1;
    for (long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x; px < hw; px += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(px / W), x = (int)(px - (long long)y * W);
        const long long pm = (long long)y * W + (W - 1 - x);          // the same pixel in the flipped image's output
        const long long g = ldg_stream_l1(gt_n + px);
        int arg;
        if (CT > 0) {
            float za[CR], zb[CR];
#pragma unroll
            for (int c = 0; c < CT; ++c) za[c] = ldg_stream_f1(a_n + (long long)c * hw + px);
#pragma unroll
            for (int c = 0; c < CT; ++c) zb[c] = ldg_stream_f1(b_n + (long long)c * hw + pm);
            float ma = za[0], mb = zb[0];
#pragma unroll
            for (int c = 1; c < CT; ++c) { ma = fmaxf(ma, za[c]); mb = fmaxf(mb, zb[c]); }
            float ea[CR], eb[CR], sa = 0.f, sb = 0.f;
            const float l2e = 1.4426950408889634f;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                ea[c] = ex2_approx((za[c] - ma) * l2e);
                eb[c] = ex2_approx((zb[c] - mb) * l2e);
                sa += ea[c];
                sb += eb[c];
            }
            const float ia = rcp_approx(sa), ib = rcp_approx(sb);
            float best = -1.f, second = -1.f;
            arg = 0;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float pc = fmaf(ea[c], ia, eb[c] * ib);
                if (pc > best) { second = best; best = pc; arg = c; }
                else second = fmaxf(second, pc);
            }
            if (!(second < best - best * 1.0e-5f)) arg = ensemble_exact(za, zb, CT, ma, mb);     // also taken for NaN
        } else {
            // generic class count: two passes over the (L1/L2-resident) columns
            float ma = -INFINITY, mb = -INFINITY;
            for (int c = 0; c < C; ++c) {
                ma = fmaxf(ma, __ldg(a_n + (long long)c * hw + px));
                mb = fmaxf(mb, __ldg(b_n + (long long)c * hw + pm));
            }
            float sa = 0.f, sb = 0.f;
            for (int c = 0; c < C; ++c) {
                sa += expf(__ldg(a_n + (long long)c * hw + px) - ma);
                sb += expf(__ldg(b_n + (long long)c * hw + pm) - mb);
            }
            float best = -1.f;
            arg = 0;
            for (int c = 0; c < C; ++c) {
                const float pa = __fdiv_rn(expf(__ldg(a_n + (long long)c * hw + px) - ma), sa);
                const float pb = __fdiv_rn(expf(__ldg(b_n + (long long)c * hw + pm) - mb), sb);
                const float pc = __fmul_rn(__fadd_rn(pa, pb), 0.5f);
                if (pc > best) { best = pc; arg = c; }
            }
        }
        if (g >= 0 && g < C) atomicAdd(&s_cm[(int)g * C + arg], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
        const unsigned v = s_cm[b];
        if (v) atomicAdd(&cm[b], (unsigned long long)v);
    }
}

// Two pixels per thread (W even, 8 / 16-byte aligned rows): the direct pair [x, x+1] and the mirrored pair
// [W-2-x, W-1-x] are one 64-bit load each per class (half the load instructions, twice the bytes in flight per
// thread), the exponentials overwrite the logits in place (no second register array) and the rare near-tie pixel
// re-reads its 2 C logits from L2 for the exact replay.  One persistent wave of 2 CTAs per SM.
__device__ __forceinline__ float2 ldg_stream_f2(const float* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}

__device__ __noinline__ int ensemble_exact_gmem(const float* pa, const float* pb, long long hw, int C, float ma, float mb) {
    float sa = 0.f, sb = 0.f;
    for (int c = 0; c < C; ++c) { sa += expf(__ldg(pa + (long long)c * hw) - ma); sb += expf(__ldg(pb + (long long)c * hw) - mb); }
    float best = -1.f;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
        const float qa = __fdiv_rn(expf(__ldg(pa + (long long)c * hw) - ma), sa);
        const float qb = __fdiv_rn(expf(__ldg(pb + (long long)c * hw) - mb), sb);
        const float pc = __fmul_rn(__fadd_rn(qa, qb), 0.5f);
        if (pc > best) { best = pc; arg = c; }
    }
    return arg;
}

template <int CT>
__global__ void __launch_bounds__(kFlipThreads, 2)
confusion_flip2_kernel(const int64_t* __restrict__ gt, const float* __restrict__ la, const float* __restrict__ lb, int C,
                       int H, int W, unsigned long long* __restrict__ cm) {
    extern __shared__ unsigned s_cm[];
    const int nbins = C * C;
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) s_cm[b] = 0u;
    __syncthreads();
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y;
    const long long hw = (long long)H * W;
    const int64_t* gt_n = gt + (long long)n * hw;
    const float* a_n = la + (long long)n * C * hw;
    const float* b_n = lb + (long long)n * C * hw;
    const int W2 = W >> 1;
    const long long npairs = (long long)H * W2;
    const float l2e = 1.4426950408889634f;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < npairs; q += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(q / W2), x = 2 * (int)(q - (long long)y * W2);
        const long long px = (long long)y * W + x;
        const long long pm = (long long)y * W + (W - 2 - x);          // mirrored pair: .y is pixel x's mirror, .x is pixel x+1's
        const longlong2 g = ldg_stream_l2(gt_n + px);
        float2 za[CT], zb[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) za[c] = ldg_stream_f2(a_n + (long long)c * hw + px);
#pragma unroll
        for (int c = 0; c < CT; ++c) zb[c] = ldg_stream_f2(b_n + (long long)c * hw + pm);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            // pixel x + j: direct logits za[c].{x,y}[j], mirrored logits zb[c].{y,x}[j]
            float ma = j ? za[0].y : za[0].x, mb = j ? zb[0].x : zb[0].y;
#pragma unroll
            for (int c = 1; c < CT; ++c) { ma = fmaxf(ma, j ? za[c].y : za[c].x); mb = fmaxf(mb, j ? zb[c].x : zb[c].y); }
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c) {          // exponentials replace the logits in place
                float& ra = j ? za[c].y : za[c].x;
                float& rb = j ? zb[c].x : zb[c].y;
                ra = ex2_approx((ra - ma) * l2e);
                rb = ex2_approx((rb - mb) * l2e);
                sa += ra;
                sb += rb;
            }
            const float ia = rcp_approx(sa), ib = rcp_approx(sb);
            float best = -1.f, second = -1.f;
            int arg = 0;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const float pc = fmaf(j ? za[c].y : za[c].x, ia, (j ? zb[c].x : zb[c].y) * ib);
                if (pc > best) { second = best; best = pc; arg = c; }
                else second = fmaxf(second, pc);
            }
            if (!(second < best - best * 1.0e-5f))     // near tie (or NaN): torch's arithmetic decides
                arg = ensemble_exact_gmem(a_n + px + j, b_n + pm + (1 - j), hw, CT, ma, mb);
            const long long gj = j ? g.y : g.x;
            if (gj >= 0 && gj < C) atomicAdd(&s_cm[(int)gj * C + arg], 1u);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
        const unsigned v = s_cm[b];
        if (v) atomicAdd(&cm[b], (unsigned long long)v);
    }
}

// The two-pixel kernel is used wherever the geometry allows it (41.0 -> 31.0 us, 63 -> 83 % of HBM on the bench shape);
// MSQ_FLIP_PX=1 forces the one-pixel kernel (A/B knob, scripts/ab_flip.py).
static int flip_px() {
    static int v = [] { const char* e = getenv("MSQ_FLIP_PX"); return (e && e[0] == '1') ? 1 : 2; }();
    return v;
}

template <int CT>
static int launch_flip(const int64_t* gt, const float* la, const float* lb, int n, int C, int H, int W,
                       unsigned long long* cm, cudaStream_t st) {
    const long long hw = (long long)H * W;
    if (CT > 0 && flip_px() == 2 && (W & 1) == 0 && ((((uintptr_t)la) | ((uintptr_t)lb)) & 7u) == 0 && (((uintptr_t)gt) & 15u) == 0) {
        const long long npairs = hw / 2;
        long long b2 = (npairs + kFlipThreads - 1) / kFlipThreads;
        const long long cap2 = ((long long)sm_count() * 2 + n - 1) / n;          // one persistent wave over all images
        if (b2 > cap2) b2 = cap2;
        if (b2 < 1) b2 = 1;
        const cudaError_t le2 = launch_pdl(confusion_flip2_kernel<(CT > 0 ? CT : 1)>, dim3((unsigned)b2, (unsigned)n), dim3(kFlipThreads),
                                           (size_t)C * C * sizeof(unsigned), st, gt, la, lb, C, H, W, cm);
        if (le2 != cudaSuccess) return (int)le2;
        MSQ_CHECK_LAUNCH();
        return 0;
    }
    long long bx = (hw + kFlipThreads - 1) / kFlipThreads;
    const long long cap = ((long long)sm_count() * 2 * 4 + n - 1) / n;      // a few waves over all images
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    const cudaError_t le = launch_pdl(confusion_flip_kernel<CT>, dim3((unsigned)bx, (unsigned)n), dim3(kFlipThreads),
                                      (size_t)C * C * sizeof(unsigned), st, gt, la, lb, C, H, W, cm);
    if (le != cudaSuccess) return (int)le;
    MSQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace msq

using namespace msq;

extern "C" int msq_confusion_flip_f32(const int64_t* gt, const float* logits, const float* logits_flipped, int n,
                                      int num_class, int out_h, int out_w, unsigned long long* cm, msq_stream_t stream) {
    if (!cm || num_class < 1 || num_class > MSQ_MAX_CLASSES || n < 0 || out_h < 0 || out_w < 0) return MSQ_E_BADARG;
    if (n == 0 || out_h == 0 || out_w == 0) return 0;
    if (!gt || !logits || !logits_flipped) return MSQ_E_BADARG;
    if (((((uintptr_t)gt) | ((uintptr_t)cm)) & 7u) || ((((uintptr_t)logits) | ((uintptr_t)logits_flipped)) & 3u)) return MSQ_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    switch (num_class) {
        case 13: return launch_flip<13>(gt, logits, logits_flipped, n, num_class, out_h, out_w, cm, st);
        case 16: return launch_flip<16>(gt, logits, logits_flipped, n, num_class, out_h, out_w, cm, st);
        case 19: return launch_flip<19>(gt, logits, logits_flipped, n, num_class, out_h, out_w, cm, st);
        default: return launch_flip<0>(gt, logits, logits_flipped, n, num_class, out_h, out_w, cm, st);
    }
}
