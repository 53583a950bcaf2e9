// Strict drop-in kernels: MaxSquareloss.forward / IW_MaxSquareloss.forward on the
// full-resolution probabilities the reference's trainers pass in
// (utils/loss.py:76-102, 110-119) and the matching dL/dprob.
//
// Forward reads prob once (4C B/pixel); backward reads it once and writes the
// gradient once (8C B/pixel): both are pure HBM streams, 128-bit loads with C
// independent requests in flight per thread.  The forward never forms the
// per-pixel weight map of utils/loss.py:96-98: it buckets q = sum_c p_c^2 by the
// pixel's argmax class (S_nk) next to the class histogram, and the last CTA
// turns hist -> weights -> loss = -(1/(N C)) sum_nk w_nk S_nk.
#include "common.cuh"

namespace msq {

constexpr int kProbThreads = 256;

// torch.max(prob, 1) semantics (utils/loss.py:84): first maximum wins; NaN is the maximum.
__device__ __forceinline__ void max_step(float v, int c, float& best, int& arg) {
    if (v > best || (v != v && best == best)) { best = v; arg = c; }
}

constexpr unsigned long long kBktMask = (1ull << 48) - 1ull;

// ------------------------------------------------------------------ K3: forward
// grid (bx, N).
// IW: every thread owns, per class, a private packed accumulator in shared memory
//     (pixel count << 48 | sum of q in 2^-32 fixed point).  A run of equal argmax
//     classes costs one conflict-free LDS.64/STS.64 pair -- no shared atomics (64-bit
//     shared atomics are CAS loops on this chip) and no dependence on how coherent the
//     input is.  The CTA reduces its buckets with warp shuffles and issues one global
//     atomic per class and warp.
// MaxSquare: one bucket; q and the count of elements != ignore (utils/loss.py:117) are
//     accumulated in registers, exactly (every 4-pixel sum is converted to fixed point).
template <int CT, bool IW, bool HAS_LABEL, bool VEC>
__global__ void __launch_bounds__(kProbThreads, CT > 0 ? 2 : 4)
prob_fwd_kernel(const float* __restrict__ prob, int n_img, int C, long long hw, const int64_t* __restrict__ label,
                float ignore_val, State st) {
    extern __shared__ __align__(16) unsigned long long s_bkt[];      // [C][kProbThreads]  (IW only)
    __shared__ unsigned s_lab[MSQ_MAX_CLASSES];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    pdl_trigger();          // the finalisation kernel may be scheduled as soon as SMs free up
    if (IW) {
        for (int c = 0; c < C; ++c) s_bkt[c * kProbThreads + tid] = 0ull;
        if (tid < MSQ_MAX_CLASSES) s_lab[tid] = 0u;
        if (HAS_LABEL) __syncthreads();
    }
    pdl_wait();             // programmatic dependent launch: global memory is touched only from here on

    const int n = blockIdx.y;
    const int rep_off = (int)(blockIdx.x % kRep) * n_img * C;        // this CTA's accumulator replica
    const float* p_n = prob + (long long)n * C * hw;
    const int64_t* lab_n = HAS_LABEL ? label + (long long)n * hw : nullptr;
    constexpr int PX = VEC ? 4 : 1;
    const long long ngroups = (hw + PX - 1) / PX;
    int run_k = -1;
    unsigned run_cnt = 0u;
    float run_q = 0.f;
    unsigned long long ms_sum = 0ull, kept = 0ull;
    bool bad = false;
    auto flush = [&]() {
        if (run_k >= 0 && run_cnt) {
            bad |= !(fabsf(run_q) < 3.0e38f);
            s_bkt[run_k * kProbThreads + tid] += to_fix(run_q) + (HAS_LABEL ? 0ull : ((unsigned long long)run_cnt << 48));
        }
    };

    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < ngroups; i += (long long)gridDim.x * blockDim.x) {
        const long long px = i * PX;
        float best[PX], q[PX];
        int arg[PX];
        unsigned kcnt[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) { arg[j] = 0; q[j] = 0.f; kcnt[j] = 0u; }
        auto visit = [&](int c, const float* v) {
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                if (c == 0) best[j] = v[j]; else max_step(v[j], c, best[j], arg[j]);
                if (IW) {
                    q[j] = fmaf(v[j], v[j], q[j]);
                } else {
                    const bool keep = (v[j] != ignore_val);          // utils/loss.py:117
                    q[j] = keep ? fmaf(v[j], v[j], q[j]) : q[j];
                    kcnt[j] += keep ? 1u : 0u;
                }
            }
        };
        if constexpr (VEC && CT > 0) {
            float4 v[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) v[c] = ldg_stream_f4(p_n + (long long)c * hw + px);
#pragma unroll
            for (int c = 0; c < CT; ++c) { const float t[4] = {v[c].x, v[c].y, v[c].z, v[c].w}; visit(c, t); }
        } else if constexpr (VEC) {
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                const float4 v = ldg_stream_f4(p_n + (long long)c * hw + px);
                const float t[4] = {v.x, v.y, v.z, v.w};
                visit(c, t);
            }
        } else {
#pragma unroll 4
            for (int c = 0; c < C; ++c) { const float t[1] = {ldg_stream_f1(p_n + (long long)c * hw + px)}; visit(c, t); }
        }
        if (IW) {
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                // utils/loss.py:85-86: a pixel whose max prob equals ignore_index is masked out
                const int k = (best[j] != ignore_val) ? arg[j] : -1;
                if (HAS_LABEL) {                                      // utils/loss.py:87-94: count `label`
                    const long long lv = lab_n[px + j];
                    if (lv >= 0 && lv < C) atomicAdd(&s_lab[(int)lv], 1u);
                }
                if (k == run_k) { run_cnt++; run_q += q[j]; }
                else { flush(); run_k = k; run_cnt = 1u; run_q = q[j]; }
            }
        } else {
            float qs = 0.f;
#pragma unroll
            for (int j = 0; j < PX; ++j) { qs += q[j]; kept += kcnt[j]; }
            bad |= !(fabsf(qs) < 3.0e38f);
            ms_sum += to_fix(qs);
        }
    }
    if (IW) {
        flush();
        __syncthreads();
        for (int c = wid; c < C; c += kProbThreads / 32) {
            unsigned cnt = 0u;
            unsigned long long sum = 0ull;
#pragma unroll
            for (int t = 0; t < kProbThreads / 32; ++t) {
                const unsigned long long v = s_bkt[c * kProbThreads + t * 32 + lane];
                cnt += (unsigned)(v >> 48);
                sum += v & kBktMask;
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            sum = warp_sum_u64(sum);
            if (lane == 0) {
                if (HAS_LABEL) cnt = s_lab[c];
                if (cnt) atomicAdd(&st.hist[rep_off + n * C + c], cnt);
                if (sum) atomicAdd(&st.sumsq[rep_off + n * C + c], sum);
            }
        }
    } else {
        ms_sum = warp_sum_u64(ms_sum);
        kept = warp_sum_u64(kept);
        if (lane == 0) {
            if (ms_sum) atomicAdd(&st.sumsq[rep_off + n * C], ms_sum);
            if (kept) atomicAdd(st.kept, kept);
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(st.flags, kFlagNonFinite);
}

// ------------------------------------------------------------------ K4: backward
// IW:        dL/dp_c = -2 * w[n, argmax] * p_c / (Nn*C) * grad_out     (0 where the pixel is masked)
// MaxSquare: dL/dp_c = -p_c / kept_global * grad_out                   (0 where p_c == ignore)
template <int CT, bool IW, bool VEC>
__global__ void __launch_bounds__(kProbThreads, CT > 0 ? 2 : 4)
prob_bwd_kernel(const float* __restrict__ prob, int C, long long hw, float ignore_val, int n_img, int n_norm,
                const float* __restrict__ weights, const unsigned long long* __restrict__ kept_out,
                const float* __restrict__ grad_out, float* __restrict__ grad) {
    __shared__ float s_w[MSQ_MAX_CLASSES];
    const int n = blockIdx.y;
    const int tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();             // programmatic dependent launch: only the launch itself overlaps the previous kernel
    const float go = *grad_out;
    float coef;
    if (IW) {
        if (tid < C) s_w[tid] = weights[n * C + tid];
        __syncthreads();
        coef = (float)(-2.0 * (double)go / ((double)n_norm * (double)C));
    } else {
        const double kept_global = (double)(*kept_out) * ((double)n_norm / (double)n_img);
        coef = (float)(-(double)go / kept_global);
    }
    const float* p_n = prob + (long long)n * C * hw;
    float* g_n = grad + (long long)n * C * hw;
    constexpr int PX = VEC ? 4 : 1;
    const long long ngroups = (hw + PX - 1) / PX;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < ngroups; i += (long long)gridDim.x * blockDim.x) {
        const long long px = i * PX;
        if constexpr (VEC && CT > 0) {
            float4 v[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) v[c] = ldg_stream_f4(p_n + (long long)c * hw + px);
            float s[4] = {coef, coef, coef, coef};
            if (IW) {
                float best[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
                int arg[4] = {0, 0, 0, 0};
#pragma unroll
                for (int c = 1; c < CT; ++c) {
                    max_step(v[c].x, c, best[0], arg[0]); max_step(v[c].y, c, best[1], arg[1]);
                    max_step(v[c].z, c, best[2], arg[2]); max_step(v[c].w, c, best[3], arg[3]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) s[j] = (best[j] != ignore_val) ? coef * s_w[arg[j]] : 0.f;
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                float4 o;
                if (IW) { o.x = s[0] * v[c].x; o.y = s[1] * v[c].y; o.z = s[2] * v[c].z; o.w = s[3] * v[c].w; }
                else {
                    o.x = (v[c].x != ignore_val) ? coef * v[c].x : 0.f; o.y = (v[c].y != ignore_val) ? coef * v[c].y : 0.f;
                    o.z = (v[c].z != ignore_val) ? coef * v[c].z : 0.f; o.w = (v[c].w != ignore_val) ? coef * v[c].w : 0.f;
                }
                stg_stream_f4(g_n + (long long)c * hw + px, o);
            }
        } else {
            // generic class count / ragged geometry: two passes over the (L2-resident) pixel group
            float s[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) s[j] = coef;
            if (IW) {
                float best[PX];
                int arg[PX];
                for (int c = 0; c < C; ++c) {
#pragma unroll
                    for (int j = 0; j < PX; ++j) {
                        const float v = __ldg(p_n + (long long)c * hw + px + j);
                        if (c == 0) { best[j] = v; arg[j] = 0; } else max_step(v, c, best[j], arg[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < PX; ++j) s[j] = (best[j] != ignore_val) ? coef * s_w[arg[j]] : 0.f;
            }
            for (int c = 0; c < C; ++c) {
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const float v = __ldg(p_n + (long long)c * hw + px + j);
                    g_n[(long long)c * hw + px + j] = IW ? s[j] * v : ((v != ignore_val) ? coef * v : 0.f);
                }
            }
        }
    }
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

int g_prob_waves = 1;     // tuning knob: grid = waves x (SMs x CTAs/SM), split over the images

static dim3 stream_grid(long long hw, int n, int px, int ctas_per_sm) {
    // persistent grid-stride CTAs: exactly `waves` x the co-resident capacity, so the per-CTA
    // epilogue (bucket reduction, global atomics) is amortised over many pixel groups
    const long long groups = (hw + px - 1) / px;
    long long bx = (groups + kProbThreads - 1) / kProbThreads;
    const long long cap = ((long long)sm_count() * ctas_per_sm * (g_prob_waves > 0 ? g_prob_waves : 1) + n - 1) / n;
    if (bx > cap) bx = cap;
    const long long need = (groups + (long long)kProbThreads * 8192 - 1) / ((long long)kProbThreads * 8192);
    if (bx < need) bx = need;          // packed 16-bit per-thread pixel counts: <= 8192 groups per thread
    if (bx < 1) bx = 1;
    return dim3((unsigned)bx, (unsigned)n);
}

template <int CT, bool IW, bool HAS_LABEL>
static int launch_fwd(const float* prob, int n, int C, long long hw, const int64_t* label, float ign, float r32,
                      float omr32, int n_norm, State st, cudaStream_t s) {
    const bool vec = ((hw & 3) == 0) && aligned16(prob);
    const size_t smem = IW ? (size_t)C * kProbThreads * sizeof(unsigned long long) : 0;
    if (vec) {
        auto k = prob_fwd_kernel<CT, IW, HAS_LABEL, true>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const cudaError_t le = launch_pdl(k, stream_grid(hw, n, 4, CT > 0 ? 2 : 4), dim3(kProbThreads), smem, s, prob, n, C,
                                          (long long)hw, label, ign, st);
        if (le != cudaSuccess) return (int)le;
    } else {
        auto k = prob_fwd_kernel<0, IW, HAS_LABEL, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const cudaError_t le = launch_pdl(k, stream_grid(hw, n, 1, 4), dim3(kProbThreads), smem, s, prob, n, C, (long long)hw,
                                          label, ign, st);
        if (le != cudaSuccess) return (int)le;
    }
    MSQ_CHECK_LAUNCH();
    return launch_finalize(st, IW ? MSQ_MODE_IW : MSQ_MODE_MAXSQUARE, n, C, r32, omr32, n_norm, 0ull, s);
}

template <int CT, bool IW>
static int launch_bwd(const float* prob, int n, int C, long long hw, float ign, int n_norm, State st,
                      const float* grad_out, float* grad, cudaStream_t s) {
    const bool vec = ((hw & 3) == 0) && aligned16(prob) && aligned16(grad);
    if (vec && CT > 0) {
        const cudaError_t le = launch_pdl(prob_bwd_kernel<CT, IW, true>, stream_grid(hw, n, 4, 2), dim3(kProbThreads), 0, s, prob, C,
                                          (long long)hw, ign, n, n_norm, (const float*)st.weights,
                                          (const unsigned long long*)st.kept_out, grad_out, grad);
        if (le != cudaSuccess) return (int)le;
    } else {
        const cudaError_t le = launch_pdl(prob_bwd_kernel<0, IW, false>, stream_grid(hw, n, 1, 4), dim3(kProbThreads), 0, s, prob, C,
                                          (long long)hw, ign, n, n_norm, (const float*)st.weights,
                                          (const unsigned long long*)st.kept_out, grad_out, grad);
        if (le != cudaSuccess) return (int)le;
    }
    MSQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace msq

using namespace msq;

extern "C" int msq_prob_fwd(int mode, const float* prob, int n, int num_class, int64_t hw, const int64_t* label,
                            double ratio, int ignore_index, int n_images_norm, void* accum, void* out, msq_stream_t stream) {
    if (!prob || !accum || !out || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES || hw < 1) return MSQ_E_BADARG;
    if (mode != MSQ_MODE_IW && mode != MSQ_MODE_MAXSQUARE) return MSQ_E_BADARG;
    if ((((uintptr_t)prob) & 3u) || ((((uintptr_t)accum) | ((uintptr_t)out)) & 15u) || (label && (((uintptr_t)label) & 7u))) return MSQ_E_ALIGN;
    const State st = carve(accum, out, n, num_class);
    const float r32 = (float)ratio, omr32 = (float)(1.0 - ratio);
    const float ign = (float)ignore_index;
    const int nn = n_images_norm > 0 ? n_images_norm : n;
    cudaStream_t s = (cudaStream_t)stream;
    const int C = num_class;
#define MSQ_FWD(CT)                                                                                              \
    (mode == MSQ_MODE_MAXSQUARE ? launch_fwd<CT, false, false>(prob, n, C, hw, nullptr, ign, r32, omr32, nn, st, s) \
     : label ? launch_fwd<CT, true, true>(prob, n, C, hw, label, ign, r32, omr32, nn, st, s)                     \
             : launch_fwd<CT, true, false>(prob, n, C, hw, nullptr, ign, r32, omr32, nn, st, s))
    switch (C) {
        case 13: return MSQ_FWD(13);
        case 16: return MSQ_FWD(16);
        case 19: return MSQ_FWD(19);
        default: return MSQ_FWD(0);
    }
#undef MSQ_FWD
}

extern "C" int msq_prob_bwd(int mode, const float* prob, int n, int num_class, int64_t hw, int ignore_index,
                            int n_images_norm, const void* out, const float* grad_out, float* grad_prob,
                            msq_stream_t stream) {
    if (!prob || !out || !grad_out || !grad_prob || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES || hw < 1)
        return MSQ_E_BADARG;
    if (mode != MSQ_MODE_IW && mode != MSQ_MODE_MAXSQUARE) return MSQ_E_BADARG;
    if ((((uintptr_t)prob) | ((uintptr_t)grad_prob) | ((uintptr_t)grad_out)) & 3u) return MSQ_E_ALIGN;
    const State st = carve(nullptr, const_cast<void*>(out), n, num_class);
    const float ign = (float)ignore_index;
    const int nn = n_images_norm > 0 ? n_images_norm : n;
    cudaStream_t s = (cudaStream_t)stream;
    const int C = num_class;
#define MSQ_BWD(CT)                                                                              \
    (mode == MSQ_MODE_IW ? launch_bwd<CT, true>(prob, n, C, hw, ign, nn, st, grad_out, grad_prob, s) \
                         : launch_bwd<CT, false>(prob, n, C, hw, ign, nn, st, grad_out, grad_prob, s))
    switch (C) {
        case 13: return MSQ_BWD(13);
        case 16: return MSQ_BWD(16);
        case 19: return MSQ_BWD(19);
        default: return MSQ_BWD(0);
    }
#undef MSQ_BWD
}
