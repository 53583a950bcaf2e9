// Strict MinEnt kernels: softCrossEntropy.forward / IWsoftCrossEntropy.forward on full-resolution `inputs` AND an
// arbitrary `target` distribution, exactly as the reference's classes define them (utils/loss.py:17-35, 37-67):
//     mask = target != ignore_index
//     ll   = log_softmax(inputs, 1)
//     softCrossEntropy:    mean( (-ll * target)[mask] )
//     IWsoftCrossEntropy:  sum( (-ll * target * w[argmax_c inputs])[mask] ) / (N C),   w from the per-image histc of
//                          argmax(inputs) (bins=C, min=0, max=C-1: a bincount), w_k = 1/max(hist_k^r total^(1-r), 1)
// The trainers pass target = softmax(inputs) (tools/solve_gta5.py:188-190,199), for which the fused kernels of
// fused_loss.cu exist; these kernels honour ANY target.  Both tensors are streamed once (8C B/pixel forward;
// backward reads 8C and writes 4C or 8C B/pixel): HBM-bound, two pixels per thread, 2C independent 64-bit loads in
// flight per thread.  The backward returns d/d inputs and, when asked, d/d target (the trainers' target is attached
// to the graph):
//     dL/dz_j = a (p_j T - t_j [t_j kept]),   T = sum over kept c of t_c
//     dL/dt_j = a (lse - z_j) [t_j kept],     a = grad_out / kept   |   grad_out w[k] / (N C)
#include "common.cuh"

namespace msq {

constexpr int kSceThreads = 256;

__device__ __forceinline__ float2 ldg_stream_f2(const float* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f2(float* p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// torch.max(inputs, 1) (utils/loss.py:54): first maximum wins; NaN is the maximum.
__device__ __forceinline__ void sce_max_step(float v, int c, float& best, int& arg) {
    if (v > best || (v != v && best == best)) { best = v; arg = c; }
}

// per pixel, from the C logits z and targets t of ONE pixel held in registers / re-read by `get`
struct ScePixel {
    float lse;      // log of s = sum_c exp(z_c - m):  -log_softmax_c = lse - (z_c - m), formed like torch's kernel does
                    // ((z - max) - log(sum)), so that the dominant class, whose z - m is exactly 0, keeps full precision
    float m, inv_s;
    float E;        // sum over kept c of t_c (lse - (z_c - m))
    float T;        // sum over kept c of t_c
    unsigned kept;
    int k;
};

template <typename GetZ, typename GetT>
__device__ __forceinline__ ScePixel sce_pixel(int C, float ign, GetZ z, GetT t) {
    ScePixel r;
    float best = z(0);
    int k = 0;
    for (int c = 1; c < C; ++c) sce_max_step(z(c), c, best, k);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += __expf(z(c) - best);
    r.m = best;
    r.k = k;
    r.inv_s = __fdividef(1.0f, s);
    r.lse = logf(s);          // logf, not lg2.approx: its 2^-22 ABSOLUTE error is a bias where the entropy is ~1e-3
    float E = 0.f, T = 0.f;
    unsigned kept = 0u;
    for (int c = 0; c < C; ++c) {
        const float tc = t(c);
        const bool keep = (tc != ign);                       // utils/loss.py:30,53
        E = keep ? fmaf(tc, r.lse - (z(c) - best), E) : E;
        T = keep ? T + tc : T;
        kept += keep ? 1u : 0u;
    }
    r.E = E; r.T = T; r.kept = kept;
    return r;
}

// ------------------------------------------------------------------ forward
// grid (bx, N).  IW: per-thread private buckets in shared memory ([C][threads] fp64 sum + [C][threads] count), updated
// once per run of equal argmax classes; reduced per CTA with warp shuffles, one global atomic per class and warp.
template <int CT, bool IW, int PX>
__global__ void __launch_bounds__(kSceThreads, 2)
softce_fwd_kernel(const float* __restrict__ inputs, const float* __restrict__ target, int n_img, int C, long long hw,
                  float ign, State st) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    double* s_sum = (double*)s_raw;                                           // [C][kSceThreads]   (IW only)
    unsigned* s_cnt = (unsigned*)(s_raw + (size_t)C * kSceThreads * 8);       // [C][kSceThreads]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    pdl_trigger();
    if (IW) for (int c = 0; c < C; ++c) { s_sum[c * kSceThreads + tid] = 0.0; s_cnt[c * kSceThreads + tid] = 0u; }
    pdl_wait();
    const int n = blockIdx.y;
    const int rep_off = (int)(blockIdx.x % kRep) * n_img * C;
    const float* z_n = inputs + (long long)n * C * hw;
    const float* t_n = target + (long long)n * C * hw;
    double* acc = (double*)st.sumsq;          // this path keeps sums of E as fp64 (targets are arbitrary reals)
    const long long ngroups = (hw + PX - 1) / PX;
    int run_k = -1;
    unsigned run_cnt = 0u;
    float run_e = 0.f;
    double ms_sum = 0.0;
    unsigned long long kept = 0ull;
    bool bad = false;
    auto flush = [&]() {
        if (run_cnt) {
            bad |= !(fabsf(run_e) < 3.0e38f);
            s_sum[run_k * kSceThreads + tid] += (double)run_e;
            s_cnt[run_k * kSceThreads + tid] += run_cnt;
        }
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < ngroups; i += (long long)gridDim.x * blockDim.x) {
        const long long px = i * PX;
        ScePixel r[PX];
        if constexpr (CT > 0 && PX == 2) {
            float2 zv[CT], tv[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) zv[c] = ldg_stream_f2(z_n + (long long)c * hw + px);
#pragma unroll
            for (int c = 0; c < CT; ++c) tv[c] = ldg_stream_f2(t_n + (long long)c * hw + px);
            // fully unrolled: the lambdas index register arrays with compile-time constants
            auto run = [&](auto sel) {
                ScePixel q;
                float best = sel(zv[0]);
                int k = 0;
#pragma unroll
                for (int c = 1; c < CT; ++c) sce_max_step(sel(zv[c]), c, best, k);
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < CT; ++c) s += __expf(sel(zv[c]) - best);
                q.m = best; q.k = k; q.inv_s = __fdividef(1.0f, s); q.lse = logf(s);
                float E = 0.f, T = 0.f;
                unsigned kp = 0u;
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    const float tc = sel(tv[c]);
                    const bool keep = (tc != ign);
                    E = keep ? fmaf(tc, q.lse - (sel(zv[c]) - best), E) : E;
                    T = keep ? T + tc : T;
                    kp += keep ? 1u : 0u;
                }
                q.E = E; q.T = T; q.kept = kp;
                return q;
            };
            r[0] = run([](const float2& v) { return v.x; });
            r[1] = run([](const float2& v) { return v.y; });
        } else {
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const long long p = px + j;
                if (p < hw) {
                    r[j] = sce_pixel(C, ign, [&](int c) { return __ldg(z_n + (long long)c * hw + p); },
                                     [&](int c) { return __ldg(t_n + (long long)c * hw + p); });
                } else {
                    r[j].E = 0.f; r[j].T = 0.f; r[j].kept = 0u; r[j].k = -1; r[j].lse = 0.f; r[j].m = 0.f; r[j].inv_s = 0.f;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            if (px + j >= hw) continue;
            kept += r[j].kept;
            if (IW) {
                if (r[j].k == run_k) { run_cnt++; run_e += r[j].E; }
                else { flush(); run_k = r[j].k; run_cnt = 1u; run_e = r[j].E; }
            } else {
                bad |= !(fabsf(r[j].E) < 3.0e38f);
                ms_sum += (double)r[j].E;
            }
        }
    }
    if (IW) {
        flush();
        __syncthreads();
        for (int c = wid; c < C; c += kSceThreads / 32) {
            unsigned cnt = 0u;
            double sum = 0.0;
#pragma unroll
            for (int t = 0; t < kSceThreads / 32; ++t) {
                cnt += s_cnt[c * kSceThreads + t * 32 + lane];
                sum += s_sum[c * kSceThreads + t * 32 + lane];
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) {
                if (cnt) atomicAdd(&st.hist[rep_off + n * C + c], cnt);
                if (sum != 0.0) atomicAdd(&acc[rep_off + n * C + c], sum);
            }
        }
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ms_sum += __shfl_xor_sync(0xffffffffu, ms_sum, o);
        if (lane == 0 && ms_sum != 0.0) atomicAdd(&acc[rep_off + n * C], ms_sum);
    }
    kept = warp_sum_u64(kept);
    if (lane == 0 && kept) atomicAdd(st.kept, kept);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(st.flags, kFlagNonFinite);
}

// ------------------------------------------------------------------ backward
template <int CT, bool IW, int PX, bool WANT_DT>
__global__ void __launch_bounds__(kSceThreads, 2)
softce_bwd_kernel(const float* __restrict__ inputs, const float* __restrict__ target, int C, long long hw, float ign,
                  int n_img, int n_norm, const float* __restrict__ weights, const unsigned long long* __restrict__ kept_out,
                  const float* __restrict__ grad_out, float* __restrict__ grad_in, float* __restrict__ grad_t) {
    __shared__ float s_w[MSQ_MAX_CLASSES];
    const int n = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();
    const float go = *grad_out;
    float coef;
    if (IW) {
        if (tid < C) s_w[tid] = weights[n * C + tid];
        __syncthreads();
        coef = (float)((double)go / ((double)n_norm * (double)C));
    } else {
        const double kept_global = (double)(*kept_out) * ((double)n_norm / (double)n_img);
        coef = (float)((double)go / kept_global);
    }
    const float* z_n = inputs + (long long)n * C * hw;
    const float* t_n = target + (long long)n * C * hw;
    float* gz_n = grad_in + (long long)n * C * hw;
    float* gt_n = WANT_DT ? grad_t + (long long)n * C * hw : nullptr;
    const long long ngroups = (hw + PX - 1) / PX;
    for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < ngroups; i += (long long)gridDim.x * blockDim.x) {
        const long long px = i * PX;
        if constexpr (CT > 0 && PX == 2) {
            float2 zv[CT], tv[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) zv[c] = ldg_stream_f2(z_n + (long long)c * hw + px);
#pragma unroll
            for (int c = 0; c < CT; ++c) tv[c] = ldg_stream_f2(t_n + (long long)c * hw + px);
            float m[2] = {zv[0].x, zv[0].y};
            int k[2] = {0, 0};
#pragma unroll
            for (int c = 1; c < CT; ++c) { sce_max_step(zv[c].x, c, m[0], k[0]); sce_max_step(zv[c].y, c, m[1], k[1]); }
            float s[2] = {0.f, 0.f}, T[2] = {0.f, 0.f};
            float2 e[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                e[c] = make_float2(__expf(zv[c].x - m[0]), __expf(zv[c].y - m[1]));
                s[0] += e[c].x; s[1] += e[c].y;
                T[0] += (tv[c].x != ign) ? tv[c].x : 0.f;
                T[1] += (tv[c].y != ign) ? tv[c].y : 0.f;
            }
            const float a0 = IW ? coef * s_w[k[0]] : coef, a1 = IW ? coef * s_w[k[1]] : coef;
            const float is0 = __fdividef(1.0f, s[0]), is1 = __fdividef(1.0f, s[1]);
            const float lse0 = logf(s[0]), lse1 = logf(s[1]);
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const bool k0 = tv[c].x != ign, k1 = tv[c].y != ign;
                float2 gz;
                gz.x = a0 * (e[c].x * is0 * T[0] - (k0 ? tv[c].x : 0.f));
                gz.y = a1 * (e[c].y * is1 * T[1] - (k1 ? tv[c].y : 0.f));
                stg_stream_f2(gz_n + (long long)c * hw + px, gz);
                if (WANT_DT) {
                    float2 gt;
                    gt.x = k0 ? a0 * (lse0 - (zv[c].x - m[0])) : 0.f;
                    gt.y = k1 ? a1 * (lse1 - (zv[c].y - m[1])) : 0.f;
                    stg_stream_f2(gt_n + (long long)c * hw + px, gt);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                const long long p = px + j;
                if (p >= hw) continue;
                const ScePixel r = sce_pixel(C, ign, [&](int c) { return __ldg(z_n + (long long)c * hw + p); },
                                             [&](int c) { return __ldg(t_n + (long long)c * hw + p); });
                const float a = IW ? coef * s_w[r.k] : coef;
                for (int c = 0; c < C; ++c) {
                    const float zc = __ldg(z_n + (long long)c * hw + p), tc = __ldg(t_n + (long long)c * hw + p);
                    const bool keep = tc != ign;
                    gz_n[(long long)c * hw + p] = a * (__expf(zc - r.m) * r.inv_s * r.T - (keep ? tc : 0.f));
                    if (WANT_DT) gt_n[(long long)c * hw + p] = keep ? a * (r.lse - (zc - r.m)) : 0.f;
                }
            }
        }
    }
}

static dim3 sce_grid(long long hw, int n, int px) {
    const long long groups = (hw + px - 1) / px;
    long long bx = (groups + kSceThreads - 1) / kSceThreads;
    const long long cap = ((long long)sm_count() * 2 + n - 1) / n;          // one persistent wave, 2 CTAs per SM
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    return dim3((unsigned)bx, (unsigned)n);
}

template <int CT, bool IW>
static int launch_sce_fwd(const float* z, const float* t, int n, int C, long long hw, float ign, float r32, float omr32,
                          int n_norm, State st, cudaStream_t s) {
    const bool vec = CT > 0 && ((hw & 1) == 0) && (((((uintptr_t)z) | ((uintptr_t)t)) & 7u) == 0);
    const size_t smem = IW ? (size_t)C * kSceThreads * 12 : 0;
    cudaError_t le;
    if (vec) {
        auto k = softce_fwd_kernel<CT, IW, 2>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        le = launch_pdl(k, sce_grid(hw, n, 2), dim3(kSceThreads), smem, s, z, t, n, C, (long long)hw, ign, st);
    } else {
        auto k = softce_fwd_kernel<0, IW, 1>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        le = launch_pdl(k, sce_grid(hw, n, 1), dim3(kSceThreads), smem, s, z, t, n, C, (long long)hw, ign, st);
    }
    if (le != cudaSuccess) return (int)le;
    MSQ_CHECK_LAUNCH();
    // loss_kind 2: the sums are fp64 (not fixed point); same formulas as the fused MinEnt losses
    return launch_finalize(st, IW ? MSQ_MODE_IW : MSQ_MODE_MAXSQUARE, n, C, r32, omr32, n_norm, 0ull, s, 0, 2);
}

template <int CT, bool IW, bool WANT_DT>
static int launch_sce_bwd(const float* z, const float* t, int n, int C, long long hw, float ign, int n_norm, State st,
                          const float* go, float* gz, float* gt, cudaStream_t s) {
    const bool vec = CT > 0 && ((hw & 1) == 0) &&
                     (((((uintptr_t)z) | ((uintptr_t)t) | ((uintptr_t)gz) | ((uintptr_t)gt)) & 7u) == 0);
    cudaError_t le;
    if (vec)
        le = launch_pdl(softce_bwd_kernel<CT, IW, 2, WANT_DT>, sce_grid(hw, n, 2), dim3(kSceThreads), 0, s, z, t, C, (long long)hw,
                        ign, n, n_norm, (const float*)st.weights, (const unsigned long long*)st.kept_out, go, gz, gt);
    else
        le = launch_pdl(softce_bwd_kernel<0, IW, 1, WANT_DT>, sce_grid(hw, n, 1), dim3(kSceThreads), 0, s, z, t, C, (long long)hw,
                        ign, n, n_norm, (const float*)st.weights, (const unsigned long long*)st.kept_out, go, gz, gt);
    if (le != cudaSuccess) return (int)le;
    MSQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace msq

using namespace msq;

extern "C" int msq_softce_fwd(int mode, const float* inputs, const float* target, int n, int num_class, int64_t hw,
                              double ratio, int ignore_index, int n_images_norm, void* accum, void* out, msq_stream_t stream) {
    if (!inputs || !target || !accum || !out || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES || hw < 1) return MSQ_E_BADARG;
    if (mode != MSQ_MODE_IW && mode != MSQ_MODE_MAXSQUARE) return MSQ_E_BADARG;
    if (((((uintptr_t)inputs) | ((uintptr_t)target)) & 3u) || ((((uintptr_t)accum) | ((uintptr_t)out)) & 15u)) return MSQ_E_ALIGN;
    const State st = carve(accum, out, n, num_class);
    const float r32 = (float)ratio, omr32 = (float)(1.0 - ratio), ign = (float)ignore_index;
    const int nn = n_images_norm > 0 ? n_images_norm : n;
    cudaStream_t s = (cudaStream_t)stream;
    const int C = num_class;
#define MSQ_F(CT) (mode == MSQ_MODE_IW ? launch_sce_fwd<CT, true>(inputs, target, n, C, hw, ign, r32, omr32, nn, st, s) \
                                       : launch_sce_fwd<CT, false>(inputs, target, n, C, hw, ign, r32, omr32, nn, st, s))
    switch (C) {
        case 13: return MSQ_F(13);
        case 16: return MSQ_F(16);
        case 19: return MSQ_F(19);
        default: return MSQ_F(0);
    }
#undef MSQ_F
}

extern "C" int msq_softce_bwd(int mode, const float* inputs, const float* target, int n, int num_class, int64_t hw,
                              int ignore_index, int n_images_norm, const void* out, const float* grad_out,
                              float* grad_inputs, float* grad_target, msq_stream_t stream) {
    if (!inputs || !target || !out || !grad_out || !grad_inputs || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES || hw < 1)
        return MSQ_E_BADARG;
    if (mode != MSQ_MODE_IW && mode != MSQ_MODE_MAXSQUARE) return MSQ_E_BADARG;
    if ((((uintptr_t)inputs) | ((uintptr_t)target) | ((uintptr_t)grad_inputs) | ((uintptr_t)grad_target) | ((uintptr_t)grad_out)) & 3u)
        return MSQ_E_ALIGN;
    const State st = carve(nullptr, const_cast<void*>(out), n, num_class);
    const float ign = (float)ignore_index;
    const int nn = n_images_norm > 0 ? n_images_norm : n;
    cudaStream_t s = (cudaStream_t)stream;
    const int C = num_class;
#define MSQ_B(CT)                                                                                                               \
    (mode == MSQ_MODE_IW                                                                                                        \
         ? (grad_target ? launch_sce_bwd<CT, true, true>(inputs, target, n, C, hw, ign, nn, st, grad_out, grad_inputs, grad_target, s)   \
                        : launch_sce_bwd<CT, true, false>(inputs, target, n, C, hw, ign, nn, st, grad_out, grad_inputs, nullptr, s))     \
         : (grad_target ? launch_sce_bwd<CT, false, true>(inputs, target, n, C, hw, ign, nn, st, grad_out, grad_inputs, grad_target, s)  \
                        : launch_sce_bwd<CT, false, false>(inputs, target, n, C, hw, ign, nn, st, grad_out, grad_inputs, nullptr, s)))
    switch (C) {
        case 13: return MSQ_B(13);
        case 16: return MSQ_B(16);
        case 19: return MSQ_B(19);
        default: return MSQ_B(0);
    }
#undef MSQ_B
}
