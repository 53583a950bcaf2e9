// Multi-level self-produced guidance (BASELINE config 3, "MaxSquare+IW+Multi"), the inline trainer
// code tools/solve_gta5.py:183,192,206-215 == tools/solve_crosscity.py:235-243, fused with the
// head-1 adaptation loss and fed from the LOW-resolution outputs of both classifier heads:
//
//   P1 = softmax(up(head1)), P2 = softmax(up(head2))                      up = bilinear, align_corners
//   loss1   = MaxSquare / IW-MaxSquare(P1)                                (as msq_fused_fwd)
//   label_2 = (max P1 > thr  or  max P2 > thr) ? argmax_c (P1 + P2)/2 : -1
//   loss2   = CrossEntropyLoss(ignore_index=-1)(up(head2), label_2)       mean over label_2 != -1
//
// One forward kernel does all of it per output pixel (column walk over both heads' tiles), writes
// the float4 statistics caches of both heads, and the two backward passes are the cached
// fused_bwd_kernel (head 1: loss gradient, head 2: GUIDE = cross-entropy gradient).
//
// label_2 is an integer map and must match the reference bit for bit.  The fast path uses
// ex2.approx probabilities; whenever the decision is close (another class within 1e-5 relative of
// the best (P1+P2), or a max probability within 1e-5 of the threshold) the pixel is re-evaluated
// with torch's own arithmetic: p = expf(z - m) / sum (sequential fp32 sum in class order),
// (p1 + p2) * 0.5f, first maximum wins, strict > against the fp32 threshold.
#include "fused_common.cuh"

namespace msq {

constexpr float kCloseRel = 1.0e-5f;

template <int CT>
__device__ __noinline__ int resolve_guidance(const float* z1, const float* z2, float m1, float m2, float thr) {
    float e1[CT], e2[CT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CT; ++c) { e1[c] = expf(z1[c] - m1); s1 += e1[c]; }
#pragma unroll
    for (int c = 0; c < CT; ++c) { e2[c] = expf(z2[c] - m2); s2 += e2[c]; }
    float best = -1.f, max1 = -1.f, max2 = -1.f;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const float p1 = __fdiv_rn(e1[c], s1), p2 = __fdiv_rn(e2[c], s2);
        max1 = fmaxf(max1, p1);
        max2 = fmaxf(max2, p2);
        const float pc = __fmul_rn(__fadd_rn(p1, p2), 0.5f);
        if (pc > best) { best = pc; arg = c; }
    }
    return (max1 > thr || max2 > thr) ? arg : -1;
}

#ifndef MSQ_MULTI_MINB
#define MSQ_MULTI_MINB 2
#endif

template <int CT, bool PAD, bool IW>
__global__ void __launch_bounds__(kTW, MSQ_MULTI_MINB)
multi_fwd_kernel(const float* __restrict__ lo1, const float* __restrict__ lo2, FusedGeo g, int n_img, unsigned units,
                 float thr, State st, void* __restrict__ aux1, void* __restrict__ aux2, float* __restrict__ zero1,
                 float* __restrict__ zero2, unsigned zero_count, long long* __restrict__ label_out) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const bool use_tab = g.R <= kRowTabMax;
    float4* s_rows = (float4*)s_raw;
    __shared__ ClassAcc s_acc;                                                // IW statistics of the current image (fused_common.cuh)
    float* s_tile1 = (float*)(s_rows + (use_tab ? g.R : 0));                  // [C][nrm][ncp]
    float* s_tile2 = s_tile1 + cpd(CT) * g.nrm * g.ncp;
    const int tid = threadIdx.x, lane = tid & 31;
    const int rep = (int)(blockIdx.x % kRep), rep_off = rep * n_img * g.C;
    pdl_trigger();
    if (IW) class_acc_zero(s_acc, tid);
    pdl_wait();            // global memory is touched only from here on (see fused_fwd_kernel)
    {
        const unsigned z0 = blockIdx.x * g.zq + blockIdx.x * g.zr / gridDim.x;
        const unsigned z1 = (blockIdx.x + 1) * g.zq + (blockIdx.x + 1) * g.zr / gridDim.x;
        for (unsigned i = z0 + tid; i < z1; i += kTW) {
            if (zero1) zero1[i] = 0.f;
            if (zero2) zero2[i] = 0.f;
        }
    }
    float4* __restrict__ ax1 = (float4*)aux1;
    float4* __restrict__ ax2 = (float4*)aux2;

    unsigned u = blockIdx.x * g.uq + blockIdx.x * g.ur / gridDim.x;   // = floor(b * units / grid) without a 64-bit division
    const unsigned u_end = (blockIdx.x + 1) * g.uq + (blockIdx.x + 1) * g.ur / gridDim.x;
    const unsigned TX = (unsigned)((g.W + kTW - 1) / kTW);
    unsigned long long ms_acc = 0ull, ce_acc = 0ull;
    unsigned nvalid = 0u;
    bool bad = false;
    constexpr int CP = (CT + 1) / 2;
    while (u < u_end) {
        const unsigned col = u / (unsigned)g.H;
        const int ys = (int)(u - col * (unsigned)g.H);
        const int ye = (int)min((unsigned)g.H, (unsigned)ys + (u_end - u));
        u += (unsigned)(ye - ys);
        const Strip sp = make_strip(g, (int)col, (int)TX, ys, ye);
        __syncthreads();
        load_tile_issue<CT, PAD>(s_tile1, lo1, g, sp);
        load_tile_issue<CT, PAD>(s_tile2, lo2, g, sp);
        if (use_tab) fill_row_table(s_rows, g, sp.ys, sp.ye);
        cp_async_wait<0>();
        __syncthreads();

        const bool active = (sp.xs + tid) < sp.xe;
        const int x = active ? sp.xs + tid : sp.xe - 1;
        int x0, x1;
        float lx0, lx1;
        src_index(g.sx, x, g.w, x0, x1, lx0, lx1);
        const int j0 = x0 - sp.c_lo, j1 = x1 - sp.c_lo;

        float2 Ha1[CP], Hb1[CP], Ha2[CP], Hb2[CP];
        int ra = -1, rb = -1;
        int run_k = -1;
        unsigned run_cnt = 0u;
        float run_q = 0.f, ce_run = 0.f;
        auto flush = [&]() {
            if (run_cnt) {
                bad |= !(fabsf(run_q) < 3.0e38f);
                if (IW) class_acc_add(s_acc, run_k, to_fix(run_q), run_cnt, true);
                else ms_acc += to_fix(run_q);
            }
        };

        for (int y = sp.ys; y < sp.ye; ++y) {
            int y0, y1;
            float ly0, ly1;
            row_params(s_rows, g, use_tab, sp.ys, y, y0, y1, ly0, ly1);
            if (y0 != ra) {
                if (y0 == rb) {
#pragma unroll
                    for (int p = 0; p < CP; ++p) { Ha1[p] = Hb1[p]; Ha2[p] = Hb2[p]; }
                } else {
                    hline<CT, PAD>(Ha1, s_tile1, g, y0 - sp.r_lo, j0, j1, lx0, lx1);
                    hline<CT, PAD>(Ha2, s_tile2, g, y0 - sp.r_lo, j0, j1, lx0, lx1);
                }
                ra = y0;
            }
            if (y1 != rb) {
                if (y1 == ra) {
#pragma unroll
                    for (int p = 0; p < CP; ++p) { Hb1[p] = Ha1[p]; Hb2[p] = Ha2[p]; }
                } else {
                    hline<CT, PAD>(Hb1, s_tile1, g, y1 - sp.r_lo, j0, j1, lx0, lx1);
                    hline<CT, PAD>(Hb2, s_tile2, g, y1 - sp.r_lo, j0, j1, lx0, lx1);
                }
                rb = y1;
            }
            const float2 w0 = splat(ly0), w1 = splat(ly1);
            // ---- head 1: the adaptation loss statistics (identical to fused_fwd_kernel)
            float2 z1[CP], e1[CP];
#pragma unroll
            for (int p = 0; p < CP; ++p) z1[p] = __ffma2_rn(Ha1[p], w0, __fmul2_rn(Hb1[p], w1));
            float m1, is1, q, qs;
            const int k1 = pixel_stats<CT, IW>(z1, e1, m1, is1, q, qs);
            // ---- head 2: softmax
            float2 z2[CP], e2[CP];
#pragma unroll
            for (int p = 0; p < CP; ++p) z2[p] = __ffma2_rn(Ha2[p], w0, __fmul2_rn(Hb2[p], w1));
            float m2 = z2[0].x;
#pragma unroll
            for (int c = 1; c < CT; ++c) m2 = fmaxf(m2, lane_of(z2[c >> 1], c));
            float s2;
            {
                const float2 l2e = splat(kLog2e), nm = splat(-m2 * kLog2e);
                float2 sa = make_float2(0.f, 0.f), sb = sa;
#pragma unroll
                for (int p = 0; p < CP; ++p) {
                    const float2 t = __ffma2_rn(z2[p], l2e, nm);
                    e2[p] = ex2_pair<CT>(t, p);
                    if (p & 1) sb = __fadd2_rn(sb, e2[p]); else sa = __fadd2_rn(sa, e2[p]);
                }
                const float2 ss = __fadd2_rn(sa, sb);
                s2 = ss.x + ss.y;
            }
            const float is2 = rcp_approx(s2);
            // ---- ensemble argmax of P1 + P2 (the /2 does not change the argmax)
            float2 pc[CP];
            {
                const float2 a1 = splat(is1), a2 = splat(is2);
#pragma unroll
                for (int p = 0; p < CP; ++p) pc[p] = __ffma2_rn(e1[p], a1, __fmul2_rn(e2[p], a2));
            }
            float bv = pc[0].x;
#pragma unroll
            for (int c = 1; c < CT; ++c) bv = fmaxf(bv, lane_of(pc[c >> 1], c));
            const float near = bv - bv * kCloseRel;
            unsigned mask_a = 0u, mask_b = 0u;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                if (c & 1)
                    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}"
                        : "+r"(mask_b) : "f"(pc[c >> 1].y), "f"(near), "r"(1u << c));
                else
                    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}"
                        : "+r"(mask_a) : "f"(pc[c >> 1].x), "f"(near), "r"(1u << c));
            }
            const unsigned mask = mask_a | mask_b;
            int lab = __ffs(mask) - 1;
            const bool valid_fast = (is1 > thr) || (is2 > thr);
            const bool close = (mask & (mask - 1u)) || fabsf(is1 - thr) <= thr * kCloseRel ||
                               fabsf(is2 - thr) <= thr * kCloseRel || lab < 0;
            if (close) {                           // rare: replay torch's arithmetic for this pixel
                asm volatile("" ::: "memory");
                float zl1[CT], zl2[CT];
#pragma unroll
                for (int c = 0; c < CT; ++c) { zl1[c] = lane_of(z1[c >> 1], c); zl2[c] = lane_of(z2[c >> 1], c); }
                lab = resolve_guidance<CT>(zl1, zl2, m1, m2, thr);
            } else if (!valid_fast) {
                lab = -1;
            }
            if (active) {
                const long long px = ((long long)sp.n * g.H + y) * g.W + x;
                if (aux1) ax1[px] = make_float4(m1, qs, is1 * is1, __int_as_float(k1));
                if (aux2) ax2[px] = make_float4(m2, is2, __int_as_float(lab), 0.f);
                if (label_out) label_out[px] = (long long)lab;
                if (lab >= 0) {                    // -log softmax(z2)[lab] = m2 + ln(s2) - z2[lab]
                    // z2[lab] re-derived from the tile with the very same arithmetic (cheaper than a C-way select)
                    constexpr int CPD = cpd(CT);
                    const float* t0 = s_tile2 + ((y0 - sp.r_lo) * g.ncp) * CPD + lab;
                    const float* t1 = s_tile2 + ((y1 - sp.r_lo) * g.ncp) * CPD + lab;
                    const float ha = __fmaf_rn(t0[j0 * CPD], lx0, __fmul_rn(t0[j1 * CPD], lx1));
                    const float hb = __fmaf_rn(t1[j0 * CPD], lx0, __fmul_rn(t1[j1 * CPD], lx1));
                    const float zsel = __fmaf_rn(ha, ly0, __fmul_rn(hb, ly1));
                    ce_run += (m2 - zsel) + logf(s2);        // full-precision log: lg2.approx's 2^-22 absolute error is a bias at -log p ~ 1e-2
                    nvalid++;
                }
                if (IW) {
                    if (k1 == run_k) { run_cnt++; run_q += q; }
                    else { flush(); run_k = k1; run_cnt = 1u; run_q = q; }
                } else {
                    run_k = 0; run_cnt++; run_q += q;
                }
            }
        }
        flush();
        bad |= !(fabsf(ce_run) < 3.0e38f);
        ce_acc += to_fix(fmaxf(ce_run, 0.f));

        if (IW) {
            __syncthreads();
            if (tid < g.C) {
                unsigned cnt;
                unsigned long long sum;
                class_acc_take(s_acc, tid, cnt, sum);
                if (cnt) atomicAdd(&st.hist[rep_off + sp.n * g.C + tid], cnt);
                if (sum) atomicAdd(&st.sumsq[rep_off + sp.n * g.C + tid], sum);
            }
        } else {
            const int next_n = (u < u_end) ? (int)((u / (unsigned)g.H) / TX) : -1;
            if (next_n != sp.n) {
                ms_acc = warp_sum_u64(ms_acc);
                if (lane == 0 && ms_acc) atomicAdd(&st.sumsq[rep_off + sp.n * g.C], ms_acc);
                ms_acc = 0ull;
            }
        }
    }
    ce_acc = warp_sum_u64(ce_acc);
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    if (lane == 0) {
        if (ce_acc) atomicAdd(&st.ce[rep], ce_acc);
        if (nvalid) atomicAdd(&st.nvalid[rep], (unsigned long long)nvalid);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(st.flags, kFlagNonFinite);
}

static inline size_t multi_smem(const FusedGeo& g, bool iw, int ct) {
    (void)iw;
    return row_tab_bytes(g) + 2 * tile_bytes(g, ct);
}

template <int CT, bool PAD>
static int launch_multi_fwd(int mode, const float* lo1, const float* lo2, int C, int h, int w, int H, int W, int n, float thr,
                            float r32, float omr32, int nn, State st, void* aux1, void* aux2, float* zero1, float* zero2,
                            long long* label_out, cudaStream_t s) {
    const bool iw = mode != MSQ_MODE_MAXSQUARE;
    const unsigned zero_count = (zero1 || zero2) ? (unsigned)((size_t)n * C * h * w) : 0u;
#define MSQ_LAUNCH(K)                                                                          \
    do {                                                                                       \
        LaunchPlan lp;                                                                         \
        const int rc = plan_launch(K, C, h, w, H, W, n, MSQ_MULTI_MINB,                                   \
                                   [&](const FusedGeo& g) { return multi_smem(g, iw, CT); }, lp); \
        if (rc) return rc;                                                                     \
        const cudaError_t le = launch_pdl(K, dim3(lp.p.grid), dim3(kTW), lp.smem, s, lo1, lo2, lp.p.g, n, (unsigned)lp.p.units, \
                                          thr, st, aux1, aux2, zero1, zero2, zero_count, label_out); \
        if (le != cudaSuccess) return (int)le;                                                 \
    } while (0)
    if (iw) MSQ_LAUNCH((multi_fwd_kernel<CT, PAD, true>));
    else MSQ_LAUNCH((multi_fwd_kernel<CT, PAD, false>));
#undef MSQ_LAUNCH
    MSQ_CHECK_LAUNCH();
    return launch_finalize(st, mode, n, C, r32, omr32, nn, (unsigned long long)n * C * H * W, s, 1);
}

template <int CT, bool PAD>
static int launch_guidance_bwd(const float* lo2, int C, int h, int w, int H, int W, int n, State st, const float* grad_out,
                               float* grad_lo, const void* aux2, bool grad_is_zeroed, cudaStream_t s) {
    if (!grad_is_zeroed) {
        cudaError_t e = cudaMemsetAsync(grad_lo, 0, (size_t)n * C * h * w * sizeof(float), s);
        if (e != cudaSuccess) return (int)e;
    }
    auto K = fused_bwd_kernel<CT, PAD, false, true, true>;
    LaunchPlan lp;
    const int rc = plan_launch(K, C, h, w, H, W, n, MSQ_BWD_MINB, [&](const FusedGeo& g) { return bwd_smem(g, CT); }, lp, 1);
    if (rc) return rc;
    const cudaError_t le = launch_pdl(K, dim3(lp.p.grid), dim3(kTW), lp.smem, s, lo2, lp.p.g, n, (unsigned)lp.p.units, n,
                                      (const float*)st.weights, grad_out, 0.f, grad_lo, aux2,
                                      (const unsigned long long*)st.nvalid_out, FinArgs{});
    if (le != cudaSuccess) return (int)le;
    MSQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace msq

using namespace msq;

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

extern "C" int msq_multi_fwd(int mode, const float* logits1, const float* logits2, int n, int num_class, int h, int w,
                             int out_h, int out_w, double ratio, double threshold, int n_images_norm, void* accum,
                             void* out, void* aux1, void* aux2, float* zero_grad1, float* zero_grad2, int64_t* label2_out,
                             msq_stream_t stream) {
    if (!logits1 || !logits2 || !accum || !out || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES || h < 1 || w < 1 ||
        out_h < 1 || out_w < 1)
        return MSQ_E_BADARG;
    if (mode != MSQ_MODE_IW && mode != MSQ_MODE_MAXSQUARE) return MSQ_E_BADARG;
    if (((((uintptr_t)logits1) | ((uintptr_t)logits2) | ((uintptr_t)zero_grad1) | ((uintptr_t)zero_grad2)) & 3u) ||
        ((((uintptr_t)accum) | ((uintptr_t)out) | ((uintptr_t)aux1) | ((uintptr_t)aux2)) & 15u) || (((uintptr_t)label2_out) & 7u))
        return MSQ_E_ALIGN;
    const State st = carve(accum, out, n, num_class);
    const float r32 = (float)ratio, omr32 = (float)(1.0 - ratio), thr = (float)threshold;
    const int nn = n_images_norm > 0 ? n_images_norm : n;
    cudaStream_t s = (cudaStream_t)stream;
#define CALL(CT, PAD) launch_multi_fwd<CT, PAD>(mode, logits1, logits2, num_class, h, w, out_h, out_w, n, thr, r32, omr32, nn, st, aux1, aux2, zero_grad1, zero_grad2, (long long*)label2_out, s)
    MSQ_DISPATCH_C(num_class, CALL)
#undef CALL
}

extern "C" int msq_guidance_bwd(const float* logits2, int n, int num_class, int h, int w, int out_h, int out_w,
                                const void* out, const void* aux2, const float* grad_out, float* grad_logits2,
                                int grad_is_zeroed, msq_stream_t stream) {
    if (!logits2 || !out || !aux2 || !grad_out || !grad_logits2 || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES ||
        h < 1 || w < 1 || out_h < 1 || out_w < 1)
        return MSQ_E_BADARG;
    if (((((uintptr_t)logits2) | ((uintptr_t)grad_logits2) | ((uintptr_t)grad_out)) & 3u) || (((uintptr_t)aux2) & 15u)) return MSQ_E_ALIGN;
    const State st = carve(nullptr, const_cast<void*>(out), n, num_class);
    cudaStream_t s = (cudaStream_t)stream;
#define CALL(CT, PAD) launch_guidance_bwd<CT, PAD>(logits2, num_class, h, w, out_h, out_w, n, st, grad_out, grad_logits2, aux2, grad_is_zeroed != 0, s)
    MSQ_DISPATCH_C(num_class, CALL)
#undef CALL
}
