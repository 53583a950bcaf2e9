// Confusion-matrix kernels: Eval.__generate_matrix / add_batch
// (reference utils/eval.py:109-121) and the callers' np.argmax(pred, axis=1)
// (tools/train_source.py:282,459) fused in front of it.
//
// Both kernels are HBM-bound streaming reads (16 B/pixel for int64 gt + int64
// pred; 4C+8 B/pixel for fp32 logits + int64 gt) into a C x C histogram.  The
// histogram lives in shared memory per CTA (<= 4 KB) and is merged into the
// caller's uint64 matrix with one global atomic per non-zero bin per CTA.
// What limits a histogram on this chip is shared-memory atomic throughput
// (~2 cycles per lane-address), not HBM, so the work per pixel is organised to
// issue as few shared atomics as possible: every thread owns 4 consecutive
// pixels and run-length-merges them in registers (segmentation maps are
// piecewise constant), and a warp-level pass (AGG) then merges equal bins
// across lanes before touching shared memory.
#include "common.cuh"

namespace msq {

constexpr int kConfThreads = 256;

// bin index of one (gt, pred) pair in the reference's flattened C*gt+pred space,
// -1 when the pixel does not count.  Out-of-contract predictions are reported
// the way numpy would have failed: errs[0] negative index, errs[1] index >= C*C.
__device__ __forceinline__ int conf_bin(long long g, long long p, int C, unsigned* errs) {
    if (g < 0 || g >= C) return -1;                       // utils/eval.py:111
    const long long idx = (long long)C * g + p;          // utils/eval.py:112
    if (idx < 0) { if (errs) atomicOr(&errs[0], 1u); return -1; }
    if (idx >= (long long)C * C) { if (errs) atomicOr(&errs[1], 1u); return -1; }
    return (int)idx;
}

// Add `cnt` to bin `bin` (bin < 0: nothing) for every lane of a converged warp.
//   AGG 0: one shared atomic per lane.
//   AGG 1: lanes holding the same bin as the first active lane are summed with one
//          REDUX and added once; the rest fall back to one atomic each.
//   AGG 2: repeat the leader pass until every distinct bin has been added once.
template <int AGG>
__device__ __forceinline__ void warp_hist_add(unsigned* s_cm, int bin, unsigned cnt) {
    if (AGG == 0) {
        if (bin >= 0) atomicAdd(&s_cm[bin], cnt);
        return;
    }
    const unsigned lane = threadIdx.x & 31u;
    unsigned active = __ballot_sync(0xffffffffu, bin >= 0);
    while (active) {
        const int leader = __ffs(active) - 1;
        const int lb = __shfl_sync(0xffffffffu, bin, leader);
        const bool mine = (bin == lb);
        const unsigned same = __ballot_sync(0xffffffffu, mine);
        const unsigned tot = __reduce_add_sync(0xffffffffu, mine ? cnt : 0u);
        if (lane == (unsigned)leader) atomicAdd(&s_cm[lb], tot);
        active &= ~same;
        if (AGG == 1) {
            if (!mine && bin >= 0) atomicAdd(&s_cm[bin], cnt);
            break;
        }
    }
}

// Run-length merge of 4 consecutive bins held by one thread, then 1..4 warp adds.
template <int AGG>
__device__ __forceinline__ void add_four(unsigned* s_cm, int b0, int b1, int b2, int b3) {
    // runs: (rb[j], rc[j]); unused slots have bin -1
    int rb0 = b0, rb1 = -1, rb2 = -1, rb3 = -1;
    unsigned rc0 = 1, rc1 = 0, rc2 = 0, rc3 = 0;
    int nr = 0;   // index of the current run
    auto push = [&](int b) {
        const int cur = (nr == 0) ? rb0 : (nr == 1) ? rb1 : (nr == 2) ? rb2 : rb3;
        if (b == cur) {
            if (nr == 0) ++rc0; else if (nr == 1) ++rc1; else if (nr == 2) ++rc2; else ++rc3;
        } else {
            ++nr;
            if (nr == 1) { rb1 = b; rc1 = 1; } else if (nr == 2) { rb2 = b; rc2 = 1; } else { rb3 = b; rc3 = 1; }
        }
    };
    push(b1); push(b2); push(b3);
    warp_hist_add<AGG>(s_cm, rb0, rc0);
    if (__any_sync(0xffffffffu, rb1 >= 0)) warp_hist_add<AGG>(s_cm, rb1, rc1);
    if (__any_sync(0xffffffffu, rb2 >= 0)) warp_hist_add<AGG>(s_cm, rb2, rc2);
    if (__any_sync(0xffffffffu, rb3 >= 0)) warp_hist_add<AGG>(s_cm, rb3, rc3);
}

__device__ __forceinline__ void merge_to_global(const unsigned* s_cm, int nbins, unsigned long long* cm) {
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
        const unsigned v = s_cm[b];
        if (v) atomicAdd(&cm[b], (unsigned long long)v);
    }
}

// ------------------------------------------------------------------ K5a: int64 gt + int64 pred
// One fat CTA per SM (1024 threads): every CTA merges its C*C bins into the caller's matrix
// with global atomics, and a few hundred CTAs finishing together queue ~10 ns deep per
// CTA on each of those addresses -- so few CTAs, each with kConfSub sub-histograms
// (one per 8 warps) to keep shared-memory atomic contention at the 256-thread level.
constexpr int kConfBig = 1024;
constexpr int kConfSub = 4;

template <int AGG, bool VEC>
__global__ void __launch_bounds__(kConfBig, 1)
confusion_i64_kernel(const int64_t* __restrict__ gt, const int64_t* __restrict__ pred, long long npix, int C,
                     unsigned long long* __restrict__ cm, unsigned* __restrict__ errs) {
    extern __shared__ unsigned s_all[];
    const int nbins = C * C;
    for (int b = threadIdx.x; b < nbins * kConfSub; b += blockDim.x) s_all[b] = 0u;
    unsigned* s_cm = s_all + (threadIdx.x >> 8) * nbins;       // this warp group's sub-histogram
    __syncthreads();
    pdl_trigger();          // programmatic dependent launch: the next kernel's shared-memory set-up may overlap our tail
    pdl_wait();             // ... and ours overlapped the previous kernel's; global memory is touched from here on

    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    const long long ngroups = npix >> 2;                    // groups of 4 pixels
    // every lane of a warp runs the same number of iterations (warp-level adds inside)
    const long long warp_first = gtid - (threadIdx.x & 31);
    for (long long base = warp_first; base < ngroups; base += gstride) {
        const long long i = base + (threadIdx.x & 31);
        int b0 = -1, b1 = -1, b2 = -1, b3 = -1;
        if (i < ngroups) {
            long long g[4], p[4];
            if (VEC) {
                const longlong2 ga = ldg_stream_l2(gt + 4 * i), gb = ldg_stream_l2(gt + 4 * i + 2);
                const longlong2 pa = ldg_stream_l2(pred + 4 * i), pb = ldg_stream_l2(pred + 4 * i + 2);
                g[0] = ga.x; g[1] = ga.y; g[2] = gb.x; g[3] = gb.y;
                p[0] = pa.x; p[1] = pa.y; p[2] = pb.x; p[3] = pb.y;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { g[j] = ldg_stream_l1(gt + 4 * i + j); p[j] = ldg_stream_l1(pred + 4 * i + j); }
            }
            b0 = conf_bin(g[0], p[0], C, errs);
            b1 = conf_bin(g[1], p[1], C, errs);
            b2 = conf_bin(g[2], p[2], C, errs);
            b3 = conf_bin(g[3], p[3], C, errs);
        }
        add_four<AGG>(s_cm, b0, b1, b2, b3);
    }
    // ragged tail (npix % 4 pixels): first lanes of CTA 0
    if (blockIdx.x == 0) {
        const long long i = (ngroups << 2) + threadIdx.x;
        if (i < npix) {
            const int b = conf_bin(gt[i], pred[i], C, errs);
            if (b >= 0) atomicAdd(&s_cm[b], 1u);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
        unsigned long long v = 0ull;
#pragma unroll
        for (int k = 0; k < kConfSub; ++k) v += s_all[k * nbins + b];
        if (v) atomicAdd(&cm[b], v);
    }
}

// ------------------------------------------------------------------ K5b: argmax(logits) fused
// NaN counts as the maximum and the first maximum wins, as numpy.argmax does.
__device__ __forceinline__ void argmax_step(float v, int c, float& best, int& arg) {
    if (v > best || (v != v && best == best)) { best = v; arg = c; }
}

template <int CT, int AGG, bool VEC>
__global__ void __launch_bounds__(kConfThreads)
confusion_logits_kernel(const int64_t* __restrict__ gt, const float* __restrict__ logits, int C, long long hw,
                        unsigned long long* __restrict__ cm, long long cm_stride, unsigned long long* __restrict__ total) {
    extern __shared__ unsigned s_cm[];
    const int nbins = C * C;
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) s_cm[b] = 0u;
    __syncthreads();
    pdl_trigger();
    pdl_wait();

    const int n = blockIdx.y;
    const int64_t* gt_n = gt + (long long)n * hw;
    const float* lg_n = logits + (long long)n * C * hw;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    const long long ngroups = VEC ? (hw >> 2) : ((hw + 3) >> 2);
    const long long warp_first = gtid - (threadIdx.x & 31);
    for (long long base = warp_first; base < ngroups; base += gstride) {
        const long long i = base + (threadIdx.x & 31);
        int bins[4] = {-1, -1, -1, -1};
        if (i < ngroups) {
            const long long px = 4 * i;
            float best[4];
            int arg[4] = {0, 0, 0, 0};
            long long g[4];
            if (VEC) {
                const longlong2 ga = ldg_stream_l2(gt_n + px), gb = ldg_stream_l2(gt_n + px + 2);
                g[0] = ga.x; g[1] = ga.y; g[2] = gb.x; g[3] = gb.y;
                if (CT > 0) {
                    float4 v[CT > 0 ? CT : 1];
#pragma unroll
                    for (int c = 0; c < CT; ++c) v[c] = ldg_stream_f4(lg_n + (long long)c * hw + px);
                    best[0] = v[0].x; best[1] = v[0].y; best[2] = v[0].z; best[3] = v[0].w;
#pragma unroll
                    for (int c = 1; c < CT; ++c) {
                        argmax_step(v[c].x, c, best[0], arg[0]);
                        argmax_step(v[c].y, c, best[1], arg[1]);
                        argmax_step(v[c].z, c, best[2], arg[2]);
                        argmax_step(v[c].w, c, best[3], arg[3]);
                    }
                } else {
                    float4 v0 = ldg_stream_f4(lg_n + px);
                    best[0] = v0.x; best[1] = v0.y; best[2] = v0.z; best[3] = v0.w;
#pragma unroll 4
                    for (int c = 1; c < C; ++c) {
                        const float4 v = ldg_stream_f4(lg_n + (long long)c * hw + px);
                        argmax_step(v.x, c, best[0], arg[0]);
                        argmax_step(v.y, c, best[1], arg[1]);
                        argmax_step(v.z, c, best[2], arg[2]);
                        argmax_step(v.w, c, best[3], arg[3]);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool in = (px + j) < hw;
                    g[j] = in ? ldg_stream_l1(gt_n + px + j) : -1;
                    best[j] = in ? ldg_stream_f1(lg_n + px + j) : 0.f;
                }
                for (int c = 1; c < C; ++c) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if ((px + j) < hw) argmax_step(ldg_stream_f1(lg_n + (long long)c * hw + px + j), c, best[j], arg[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) bins[j] = (g[j] >= 0 && g[j] < C) ? (int)g[j] * C + arg[j] : -1;
        }
        add_four<AGG>(s_cm, bins[0], bins[1], bins[2], bins[3]);
    }
    if (VEC && blockIdx.x == 0) {                            // ragged tail of the image
        const long long px = (ngroups << 2) + threadIdx.x;
        if (px < hw) {
            float best = lg_n[px];
            int arg = 0;
            for (int c = 1; c < C; ++c) argmax_step(lg_n[(long long)c * hw + px], c, best, arg);
            const long long g = gt_n[px];
            if (g >= 0 && g < C) atomicAdd(&s_cm[(int)g * C + arg], 1u);
        }
    }
    __syncthreads();
    merge_to_global(s_cm, nbins, cm + (long long)n * cm_stride);       // cm_stride > 0: one matrix per image (tools/analysis.py:200-229)
    if (total) merge_to_global(s_cm, nbins, total);
}

// ------------------------------------------------------------------ K5c: K (gt, pred) pairs in ONE launch
// The reference's evaluation loops call Eval.add_batch once per image (tools/train_source.py:429-492,
// tools/evaluate.py:99-202, tools/analysis.py:200-229): one 8 MB launch each is launch/drain bound (33 % of HBM).
// Here up to kConfJobs pairs ride in the kernel parameters; the CTAs split the CONCATENATION of all pairs' 4-pixel
// groups evenly (one balanced wave whatever the image sizes) and merge their shared-memory histogram at every pair
// boundary they cross: into ONE matrix (cm_stride = 0: K deferred add_batch calls) or into one matrix per pair
// (cm_stride >= C*C: the per-image evaluation of tools/analysis.py), and optionally into a running total as well.
constexpr int kConfJobs = 32;
struct ConfJobs {
    const int64_t* gt[kConfJobs];
    const int64_t* pred[kConfJobs];
    long long npix[kConfJobs];
    long long first[kConfJobs + 1];      // first[j] = number of 4-pixel groups of pairs 0..j-1
    int k;
};

template <bool VEC>
__global__ void __launch_bounds__(kConfBig, 1)
confusion_multi_kernel(const ConfJobs jobs, int C, unsigned long long* __restrict__ cm, long long cm_stride,
                       unsigned long long* __restrict__ total, unsigned* __restrict__ errs) {
    extern __shared__ unsigned s_all[];
    const int nbins = C * C;
    for (int b = threadIdx.x; b < nbins * kConfSub; b += blockDim.x) s_all[b] = 0u;
    unsigned* s_cm = s_all + (threadIdx.x >> 8) * nbins;
    __syncthreads();
    pdl_trigger();
    pdl_wait();
    const long long G = jobs.first[jobs.k];
    const long long g0 = G * blockIdx.x / gridDim.x, g1 = G * (blockIdx.x + 1) / gridDim.x;
    int j = 0;
    while (j < jobs.k && jobs.first[j + 1] <= g0) ++j;
    for (; j < jobs.k && jobs.first[j] < g1; ++j) {
        const long long lo = max(g0, jobs.first[j]) - jobs.first[j], hi = min(g1, jobs.first[j + 1]) - jobs.first[j];
        const int64_t* __restrict__ gt = jobs.gt[j];
        const int64_t* __restrict__ pr = jobs.pred[j];
        const long long npix = jobs.npix[j];
        // every lane of a warp runs the same number of iterations: add_four votes across the whole warp
        for (long long base = lo + (threadIdx.x & ~31u); base < hi; base += blockDim.x) {
            const long long i = base + (threadIdx.x & 31u);
            long long g[4], p[4];
            int b[4];
            if (i >= hi) {
#pragma unroll
                for (int q = 0; q < 4; ++q) { g[q] = -1; p[q] = 0; }
            } else if (VEC && 4 * i + 3 < npix) {
                const longlong2 ga = ldg_stream_l2(gt + 4 * i), gb = ldg_stream_l2(gt + 4 * i + 2);
                const longlong2 pa = ldg_stream_l2(pr + 4 * i), pb = ldg_stream_l2(pr + 4 * i + 2);
                g[0] = ga.x; g[1] = ga.y; g[2] = gb.x; g[3] = gb.y;
                p[0] = pa.x; p[1] = pa.y; p[2] = pb.x; p[3] = pb.y;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool in = 4 * i + q < npix;
                    g[q] = in ? ldg_stream_l1(gt + 4 * i + q) : -1;
                    p[q] = in ? ldg_stream_l1(pr + 4 * i + q) : 0;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) b[q] = conf_bin(g[q], p[q], C, errs);
            add_four<0>(s_cm, b[0], b[1], b[2], b[3]);
        }
        const bool last = (j + 1 >= jobs.k) || (jobs.first[j + 1] >= g1);
        if (cm_stride == 0 && !last) continue;            // one matrix for all pairs: merge once, at the end
        __syncthreads();
        unsigned long long* dst = cm + (long long)j * cm_stride;
        for (int bb = threadIdx.x; bb < nbins; bb += blockDim.x) {
            unsigned long long v = 0ull;
#pragma unroll
            for (int s = 0; s < kConfSub; ++s) { v += s_all[s * nbins + bb]; s_all[s * nbins + bb] = 0u; }
            if (v) { atomicAdd(&dst[bb], v); if (total) atomicAdd(&total[bb], v); }
        }
        __syncthreads();
    }
}

int g_conf_ctas_per_sm = 1;   // tuning knob: 1024-thread CTAs per SM for the int64 kernel (1 or 2)
int g_conf_agg = 0;   // tuning knob (msq_tune_set), see api.cu; 0 measured fastest on B200
int g_conf_grid = 0;  // tuning knob: explicit CTA count of the int64 kernel (0 = automatic).  Every CTA ends with up to C*C
                      // global atomics on the SAME addresses (~10 ns each per address), so for one 8 MB image fewer CTAs
                      // may beat one per SM; for scripts/ab_conf.py sweeps

template <int AGG>
static int launch_i64(const int64_t* gt, const int64_t* pred, long long npix, int C, unsigned long long* cm,
                      unsigned* errs, cudaStream_t st) {
    const bool vec = ((((uintptr_t)gt) | ((uintptr_t)pred)) & 15u) == 0;
    const long long groups = (npix + 3) / 4;
    long long blocks = (groups + kConfBig - 1) / kConfBig;
    const long long cap = (long long)sm_count() * (g_conf_ctas_per_sm > 0 ? g_conf_ctas_per_sm : 1);
    if (blocks > cap) blocks = cap;
    if (g_conf_grid > 0 && blocks > g_conf_grid) blocks = g_conf_grid;
    if (blocks < 1) blocks = 1;
    const size_t smem = (size_t)C * C * sizeof(unsigned) * kConfSub;
    cudaError_t le;
    if (vec) le = launch_pdl(confusion_i64_kernel<AGG, true>, dim3((unsigned)blocks), dim3(kConfBig), smem, st, gt, pred, npix, C, cm, errs);
    else le = launch_pdl(confusion_i64_kernel<AGG, false>, dim3((unsigned)blocks), dim3(kConfBig), smem, st, gt, pred, npix, C, cm, errs);
    if (le != cudaSuccess) return (int)le;
    MSQ_CHECK_LAUNCH();
    return 0;
}

template <int CT, int AGG>
static int launch_logits(const int64_t* gt, const float* logits, int n, int C, long long hw, unsigned long long* cm,
                         cudaStream_t st, long long cm_stride = 0, unsigned long long* total = nullptr) {
    const bool vec = ((hw & 3) == 0) && (((uintptr_t)gt & 15u) == 0) && (((uintptr_t)logits & 15u) == 0);
    const long long groups = (hw + 3) / 4;
    long long bx = (groups + kConfThreads - 1) / kConfThreads;
    // CT>0 keeps CT float4 in registers (~100 regs): 2 CTAs of 256 threads per SM
    const long long cap = ((long long)sm_count() * (CT > 0 ? 2 : 4) * 2 + n - 1) / n;   // ~2 waves over all images
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    const dim3 grid((unsigned)bx, (unsigned)n);
    const size_t smem = (size_t)C * C * sizeof(unsigned);
    cudaError_t le;
    if (vec) le = launch_pdl(confusion_logits_kernel<CT, AGG, true>, grid, dim3(kConfThreads), smem, st, gt, logits, C, hw, cm, cm_stride, total);
    else le = launch_pdl(confusion_logits_kernel<0, AGG, false>, grid, dim3(kConfThreads), smem, st, gt, logits, C, hw, cm, cm_stride, total);
    if (le != cudaSuccess) return (int)le;
    MSQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace msq

using namespace msq;

extern "C" int msq_confusion_i64(const int64_t* gt, const int64_t* pred, int64_t npix, int num_class,
                                 unsigned long long* cm, unsigned int* errs, msq_stream_t stream) {
    if (!cm || num_class < 1 || num_class > MSQ_MAX_CLASSES || npix < 0) return MSQ_E_BADARG;
    if (npix == 0) return 0;
    if (!gt || !pred) return MSQ_E_BADARG;
    if ((((uintptr_t)gt) | ((uintptr_t)pred) | ((uintptr_t)cm)) & 7u) return MSQ_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    switch (g_conf_agg) {
        case 0: return launch_i64<0>(gt, pred, npix, num_class, cm, errs, st);
        case 2: return launch_i64<2>(gt, pred, npix, num_class, cm, errs, st);
        default: return launch_i64<1>(gt, pred, npix, num_class, cm, errs, st);
    }
}

static int confusion_logits_dispatch(const int64_t* gt, const float* logits, int n, int num_class, int64_t hw,
                                     unsigned long long* cm, long long cm_stride, unsigned long long* total, msq_stream_t stream) {
    if (!cm || num_class < 1 || num_class > MSQ_MAX_CLASSES || n < 0 || hw < 0) return MSQ_E_BADARG;
    if (n == 0 || hw == 0) return 0;
    if (!gt || !logits) return MSQ_E_BADARG;
    if ((((uintptr_t)gt) | ((uintptr_t)cm) | ((uintptr_t)total)) & 7u) return MSQ_E_ALIGN;
    if (((uintptr_t)logits) & 3u) return MSQ_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int agg = g_conf_agg;
#define MSQ_DISPATCH_CT(CT)                                                                                  \
    (agg == 0 ? launch_logits<CT, 0>(gt, logits, n, num_class, hw, cm, st, cm_stride, total)                 \
              : agg == 2 ? launch_logits<CT, 2>(gt, logits, n, num_class, hw, cm, st, cm_stride, total)      \
                         : launch_logits<CT, 1>(gt, logits, n, num_class, hw, cm, st, cm_stride, total))
    switch (num_class) {
        case 13: return MSQ_DISPATCH_CT(13);
        case 16: return MSQ_DISPATCH_CT(16);
        case 19: return MSQ_DISPATCH_CT(19);
        default: return MSQ_DISPATCH_CT(0);
    }
#undef MSQ_DISPATCH_CT
}

extern "C" int msq_confusion_logits_f32(const int64_t* gt, const float* logits, int n, int num_class, int64_t hw,
                                        unsigned long long* cm, msq_stream_t stream) {
    return confusion_logits_dispatch(gt, logits, n, num_class, hw, cm, 0, nullptr, stream);
}

extern "C" int msq_confusion_per_image_logits_f32(const int64_t* gt, const float* logits, int n, int num_class, int64_t hw,
                                                  unsigned long long* cm_per_image, unsigned long long* total,
                                                  msq_stream_t stream) {
    return confusion_logits_dispatch(gt, logits, n, num_class, hw, cm_per_image, (long long)num_class * num_class, total, stream);
}

extern "C" int msq_confusion_i64_multi(const int64_t* const* gt, const int64_t* const* pred, const int64_t* npix, int k,
                                       int num_class, unsigned long long* cm, int64_t cm_stride, unsigned long long* total,
                                       unsigned int* errs, msq_stream_t stream) {
    if (!cm || num_class < 1 || num_class > MSQ_MAX_CLASSES || k < 0) return MSQ_E_BADARG;
    if (cm_stride != 0 && cm_stride < (int64_t)num_class * num_class) return MSQ_E_BADARG;
    if (k == 0) return 0;
    if (!gt || !pred || !npix) return MSQ_E_BADARG;
    if ((((uintptr_t)cm) | ((uintptr_t)total)) & 7u) return MSQ_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)num_class * num_class * sizeof(unsigned) * kConfSub;
    for (int base = 0; base < k; base += kConfJobs) {          // kConfJobs pairs per launch (they travel as kernel parameters)
        ConfJobs jobs;
        jobs.k = 0;
        jobs.first[0] = 0;
        bool vec = true;
        for (int j = base; j < k && jobs.k < kConfJobs; ++j) {
            if (npix[j] < 0) return MSQ_E_BADARG;
            if (npix[j] == 0) continue;
            if (!gt[j] || !pred[j]) return MSQ_E_BADARG;
            if ((((uintptr_t)gt[j]) | ((uintptr_t)pred[j])) & 7u) return MSQ_E_ALIGN;
            if ((((uintptr_t)gt[j]) | ((uintptr_t)pred[j])) & 15u) vec = false;
            const int q = jobs.k++;
            // with per-pair matrices the output slot is the pair's index in the caller's table
            jobs.gt[q] = gt[j]; jobs.pred[q] = pred[j]; jobs.npix[q] = npix[j];
            jobs.first[q + 1] = jobs.first[q] + (npix[j] + 3) / 4;
            if (cm_stride != 0 && q != j - base) return MSQ_E_BADARG;      // empty pairs are not allowed in per-pair mode
        }
        if (jobs.k == 0) continue;
        long long blocks = (jobs.first[jobs.k] + kConfBig - 1) / kConfBig;
        const long long cap = (long long)sm_count() * (g_conf_ctas_per_sm > 0 ? g_conf_ctas_per_sm : 1);
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        unsigned long long* dst = cm + (long long)base * cm_stride;
        cudaError_t le;
        if (vec) le = launch_pdl(confusion_multi_kernel<true>, dim3((unsigned)blocks), dim3(kConfBig), smem, st, jobs, num_class, dst, (long long)cm_stride, total, errs);
        else le = launch_pdl(confusion_multi_kernel<false>, dim3((unsigned)blocks), dim3(kConfBig), smem, st, jobs, num_class, dst, (long long)cm_stride, total, errs);
        if (le != cudaSuccess) return (int)le;
        MSQ_CHECK_LAUNCH();
    }
    return 0;
}
