// Source-side step of every training iteration, fused from the LOW-resolution head logits:
//
//   pred     = F.interpolate(head, size, 'bilinear', align_corners=True)      graphs/models/deeplab_multi.py:124,128
//   cur_loss = nn.CrossEntropyLoss(ignore_index=-1)(pred, y)                  tools/train_source.py:128,254 (and :257 for head 2,
//                                                                             tools/solve_crosscity.py:183-190)
//   argpred  = np.argmax(pred.data.cpu().numpy(), axis=1)                     tools/train_source.py:280-282
//   Eval.add_batch(y.cpu().numpy(), argpred)                                  tools/train_source.py:283, utils/eval.py:109-121
//
// One forward kernel (column walk, as fused_fwd_kernel) computes per output pixel the interpolated
// logits, log-sum-exp, -log softmax[y], the argmax class and the confusion-matrix increment; it
// writes the float4 cache {max, 1/sum, label} that fused_bwd_kernel<GUIDE> consumes, so the backward
// is msq_guidance_bwd: dL/dlogits = grad/n_valid * (softmax - onehot(y)) through the bilinear adjoint.
// Nothing at label resolution except the int64 label map is read; nothing but the cache is written.
//
// Labels outside [0, C) are ignored by both the loss and the matrix (the datasets emit -1 for every
// void pixel; torch itself would raise a device-side assert for other out-of-range targets).
#include "fused_common.cuh"

namespace msq {

template <int CT, bool PAD>
#ifndef MSQ_SRC_MINB
#define MSQ_SRC_MINB 4
#endif
__global__ void __launch_bounds__(kTW, MSQ_SRC_MINB)
source_ce_fwd_kernel(const float* __restrict__ lo, const int64_t* __restrict__ label, FusedGeo g, int n_img, unsigned units,
                     State st, void* __restrict__ aux, float* __restrict__ zero_buf, unsigned zero_count,
                     unsigned long long* __restrict__ cm) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr int CP = (CT + 1) / 2, CPD = cpd(CT);
    const bool use_tab = g.R <= kRowTabMax;
    float4* s_rows = (float4*)s_raw;
    float* s_tile = (float*)(s_rows + (use_tab ? g.R : 0));                   // [nrm][ncp][CPD]
    unsigned* s_cm = (unsigned*)(s_tile + CPD * g.nrm * g.ncp);               // [C*C] confusion counts of this CTA
    const int tid = threadIdx.x, lane = tid & 31;
    const int rep = (int)(blockIdx.x % kRep);
    pdl_trigger();
    if (cm) for (int i = tid; i < g.C * g.C; i += kTW) s_cm[i] = 0u;
    pdl_wait();            // global memory is touched only from here on (see fused_fwd_kernel)
    if (zero_buf) {
        const unsigned z0 = blockIdx.x * g.zq + blockIdx.x * g.zr / gridDim.x;
        const unsigned z1 = (blockIdx.x + 1) * g.zq + (blockIdx.x + 1) * g.zr / gridDim.x;
        for (unsigned i = z0 + tid; i < z1; i += kTW) zero_buf[i] = 0.f;
    }
    float4* __restrict__ ax = (float4*)aux;

    unsigned u = blockIdx.x * g.uq + blockIdx.x * g.ur / gridDim.x;   // = floor(b * units / grid) without a 64-bit division
    const unsigned u_end = (blockIdx.x + 1) * g.uq + (blockIdx.x + 1) * g.ur / gridDim.x;
    const unsigned TX = (unsigned)((g.W + kTW - 1) / kTW);
    unsigned long long ce_acc = 0ull;
    unsigned nvalid = 0u;
    bool bad = false;
    while (u < u_end) {
        const unsigned col = u / (unsigned)g.H;
        const int ys = (int)(u - col * (unsigned)g.H);
        const int ye = (int)min((unsigned)g.H, (unsigned)ys + (u_end - u));
        u += (unsigned)(ye - ys);
        const Strip sp = make_strip(g, (int)col, (int)TX, ys, ye);
        __syncthreads();
        load_tile<CT, PAD>(s_tile, lo, g, sp);
        if (use_tab) fill_row_table(s_rows, g, sp.ys, sp.ye);
        __syncthreads();

        const bool active = (sp.xs + tid) < sp.xe;
        const int x = active ? sp.xs + tid : sp.xe - 1;
        int x0, x1;
        float lx0, lx1;
        src_index(g.sx, x, g.w, x0, x1, lx0, lx1);
        const int j0 = x0 - sp.c_lo, j1 = x1 - sp.c_lo;
        const long long px0 = ((long long)sp.n * g.H + sp.ys) * g.W + x;
        const int64_t* labp = label + px0;
        float4* axp = aux ? ax + px0 : nullptr;

        float2 Ha[CP], Hb[CP];
        int ra = -1, rb = -1;
        float ce_run = 0.f;
        long long lv_next = active ? ldg_stream_l1(labp) : -1;        // labels are prefetched one row ahead
        for (int y = sp.ys; y < sp.ye; ++y) {
            int y0, y1;
            float ly0, ly1;
            row_params(s_rows, g, use_tab, sp.ys, y, y0, y1, ly0, ly1);
            const long long lv = lv_next;
            if (active && y + 1 < sp.ye) { labp += g.W; lv_next = ldg_stream_l1(labp); }
            if (y0 != ra) {
                if (y0 == rb) {
#pragma unroll
                    for (int p = 0; p < CP; ++p) Ha[p] = Hb[p];
                } else {
                    hline<CT, PAD>(Ha, s_tile, g, y0 - sp.r_lo, j0, j1, lx0, lx1);
                }
                ra = y0;
            }
            if (y1 != rb) {
                if (y1 == ra) {
#pragma unroll
                    for (int p = 0; p < CP; ++p) Hb[p] = Ha[p];
                } else {
                    hline<CT, PAD>(Hb, s_tile, g, y1 - sp.r_lo, j0, j1, lx0, lx1);
                }
                rb = y1;
            }
            float2 z[CP];
            const float2 w0 = splat(ly0), w1 = splat(ly1);
#pragma unroll
            for (int p = 0; p < CP; ++p) z[p] = __ffma2_rn(Ha[p], w0, __fmul2_rn(Hb[p], w1));
            float m = z[0].x;
#pragma unroll
            for (int c = 1; c < CT; ++c) m = fmaxf(m, lane_of(z[c >> 1], c));
            // np.argmax of the interpolated logits: the first class that attains the maximum
            unsigned mask_a = 0u, mask_b = 0u;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                if (c & 1)
                    asm("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}"
                        : "+r"(mask_b) : "f"(z[c >> 1].y), "f"(m), "r"(1u << c));
                else
                    asm("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}"
                        : "+r"(mask_a) : "f"(z[c >> 1].x), "f"(m), "r"(1u << c));
            }
            int k = __ffs(mask_a | mask_b) - 1;
            if (k < 0) k = 0;
            float s;
            {
                const float2 l2e = splat(kLog2e), nm = splat(-m * kLog2e);
                float2 sa = make_float2(0.f, 0.f), sb = sa;
#pragma unroll
                for (int p = 0; p < CP; ++p) {
                    const float2 t = __ffma2_rn(z[p], l2e, nm);
                    const float2 e = ex2_pair<CT>(t, p);
                    if (p & 1) sb = __fadd2_rn(sb, e); else sa = __fadd2_rn(sa, e);
                }
                const float2 ss = __fadd2_rn(sa, sb);
                s = ss.x + ss.y;
            }
            const float is = rcp_approx(s);
            const bool valid = active && lv >= 0 && lv < (long long)g.C;
            const int lab = valid ? (int)lv : -1;
            if (active && aux) { *axp = make_float4(m, is, __int_as_float(lab), 0.f); axp += g.W; }
            if (valid) {
                // z[lab] re-derived from the tile with the very same arithmetic (cheaper than a 19-way select)
                const float* t0 = s_tile + ((y0 - sp.r_lo) * g.ncp) * CPD + lab;
                const float* t1 = s_tile + ((y1 - sp.r_lo) * g.ncp) * CPD + lab;
                const float ha = __fmaf_rn(t0[j0 * CPD], lx0, __fmul_rn(t0[j1 * CPD], lx1));
                const float hb = __fmaf_rn(t1[j0 * CPD], lx0, __fmul_rn(t1[j1 * CPD], lx1));
                const float zsel = __fmaf_rn(ha, ly0, __fmul_rn(hb, ly1));
                ce_run += (m - zsel) + logf(s);
                nvalid++;
                if (cm) atomicAdd(&s_cm[lab * g.C + k], 1u);
            }
        }
        bad |= !(fabsf(ce_run) < 3.0e38f);
        ce_acc += to_fix(fmaxf(ce_run, 0.f));
    }
    ce_acc = warp_sum_u64(ce_acc);
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    if (lane == 0) {
        if (ce_acc) atomicAdd(&st.ce[rep], ce_acc);
        if (nvalid) atomicAdd(&st.nvalid[rep], (unsigned long long)nvalid);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(st.flags, kFlagNonFinite);
    if (cm) {
        __syncthreads();
        for (int i = tid; i < g.C * g.C; i += kTW) {
            const unsigned v = s_cm[i];
            if (v) atomicAdd(&cm[i], (unsigned long long)v);
        }
    }
}

static inline size_t source_smem(const FusedGeo& g, int ct) {
    return row_tab_bytes(g) + tile_bytes(g, ct) + (size_t)g.C * g.C * sizeof(unsigned);
}

template <int CT, bool PAD>
static int launch_source_ce(const float* lo, const int64_t* label, int C, int h, int w, int H, int W, int n, State st,
                            void* aux, float* zero_buf, unsigned long long* cm, cudaStream_t s) {
    auto K = source_ce_fwd_kernel<CT, PAD>;
    LaunchPlan lp;
    const int rc = plan_launch(K, C, h, w, H, W, n, 4, [&](const FusedGeo& g) { return source_smem(g, CT); }, lp);
    if (rc) return rc;
    const unsigned zero_count = zero_buf ? (unsigned)((size_t)n * C * h * w) : 0u;
    const cudaError_t le = launch_pdl(K, dim3(lp.p.grid), dim3(kTW), lp.smem, s, lo, label, lp.p.g, n, (unsigned)lp.p.units, st,
                                      aux, zero_buf, zero_count, cm);
    if (le != cudaSuccess) return (int)le;
    MSQ_CHECK_LAUNCH();
    // finalisation: loss2 = sum / n_valid (mean over the valid pixels, 0/0 = NaN as in torch); cleans the accumulators
    return launch_finalize(st, MSQ_MODE_MAXSQUARE, n, C, 0.f, 1.f, n, (unsigned long long)n * C * H * W, s, 1);
}

}  // namespace msq

using namespace msq;

extern "C" int msq_source_ce_fwd(const float* logits, const int64_t* label, int n, int num_class, int h, int w, int out_h,
                                 int out_w, void* accum, void* out, void* aux, float* zero_grad, unsigned long long* cm,
                                 msq_stream_t stream) {
    if (!logits || !label || !accum || !out || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES || h < 1 || w < 1 ||
        out_h < 1 || out_w < 1)
        return MSQ_E_BADARG;
    if (((((uintptr_t)logits) | ((uintptr_t)zero_grad)) & 3u) || ((((uintptr_t)label) | ((uintptr_t)cm)) & 7u) ||
        ((((uintptr_t)accum) | ((uintptr_t)out) | ((uintptr_t)aux)) & 15u))
        return MSQ_E_ALIGN;
    const State st = carve(accum, out, n, num_class);
    cudaStream_t s = (cudaStream_t)stream;
#define CALL(CT, PAD) launch_source_ce<CT, PAD>(logits, label, num_class, h, w, out_h, out_w, n, st, aux, zero_grad, cm, s)
    switch (num_class) {
        case 13: return CALL(13, false);
        case 16: return CALL(16, false);
        case 19: return CALL(19, false);
        default:
            if (num_class <= 8) return CALL(8, true);
            if (num_class <= 24) return CALL(24, true);
            return CALL(32, true);
    }
#undef CALL
}
