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
#include "common.cuh"

namespace msq {

constexpr int kTW = 128;                        // output columns (= threads) per CTA
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNearTie = 2.384185791015625e-07f;   // 2^-22, see resolve_ties
constexpr float kPadLogit = -1.0e30f;           // logits of padded classes (C < CT)

// ATen/native/UpSample.h area_pixel_compute_source_index + guard_index_and_lambda
// (align_corners=True): src = scale*dst in fp32, i0 = trunc, lambda1 = src - i0.
__host__ __device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0,
                                                   float& l1) {
#ifdef __CUDA_ARCH__
    const float src = __fmul_rn(scale, (float)dst);
#else
    volatile float srcv = scale * (float)dst;
    const float src = srcv;
#endif
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    float lam = src - (float)i0;
    lam = lam < 0.f ? 0.f : (lam > 1.f ? 1.f : lam);
    l1 = lam;
    l0 = 1.0f - lam;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
}

struct FusedGeo {
    int C, h, w, H, W;
    float sy, sx;        // (in-1)/(out-1) in fp32 (0 when out == 1)
    int R;               // output rows per strip
    int nrm, ncp;        // max low-res rows / cols any strip touches (tile pitch)
};

// Exact replica of what torch's softmax + max do when two interpolated logits are
// within a few ulps: p_c = expf(z_c - m) / sum_k expf(z_k - m) in class order, then
// the FIRST class whose p equals the maximum p (= 1/sum) wins (utils/loss.py:84).
// Only called for pixels where another class is within 2^-22 of the maximum.
template <int CT>
__device__ __noinline__ int resolve_ties(const float* z, float m) {
    float e[CT];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CT; ++c) { e[c] = expf(z[c] - m); s += e[c]; }
    const float pm = __fdiv_rn(1.0f, s);
    int k = -1;
#pragma unroll
    for (int c = CT - 1; c >= 0; --c) if (__fdiv_rn(e[c], s) == pm) k = c;
    return k;
}

// Per-pixel softmax statistics from the interpolated logits z[]:
//   e[c] = 2^((z_c - m) log2 e), inv_s = 1/sum e, q = sum_c p_c^2, returns argmax class.
template <int CT, bool NEED_ARG>
__device__ __forceinline__ int pixel_stats(const float (&z)[CT], float (&e)[CT], float& inv_s, float& q) {
    float m = z[0];
#pragma unroll
    for (int c = 1; c < CT; ++c) m = fmaxf(m, z[c]);
    int k = 0;
    if (NEED_ARG) {
        const float thr = m - kNearTie;
        unsigned mask = 0u;
#pragma unroll
        for (int c = 0; c < CT; ++c) mask |= (z[c] >= thr) ? (1u << c) : 0u;
        k = __ffs(mask) - 1;
        if (mask & (mask - 1u)) {                       // more than one class within 2^-22 of the max
            float zl[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) zl[c] = z[c];
            k = resolve_ties<CT>(zl, m);
        }
        if (k < 0) k = 0;                               // NaN logits: reference yields NaN loss anyway
    }
    const float nm = -m * kLog2e;
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        e[c] = ex2_approx(fmaf(z[c], kLog2e, nm));
        s += e[c];
        ss = fmaf(e[c], e[c], ss);
    }
    inv_s = rcp_approx(s);
    q = ss * inv_s * inv_s;
    return k;
}

// Column bookkeeping shared by both kernels.
struct Strip {
    int n, xs, xe, ys, ye;        // output extent of this CTA
    int c_lo, r_lo, nc, nr;       // low-res tile origin / extent
};

__device__ __forceinline__ Strip make_strip(const FusedGeo& g) {
    Strip s;
    s.n = blockIdx.z;
    s.xs = blockIdx.x * kTW;
    s.xe = min(g.W, s.xs + kTW);
    s.ys = blockIdx.y * g.R;
    s.ye = min(g.H, s.ys + g.R);
    int i0, i1;
    float l0, l1;
    src_index(g.sx, s.xs, g.w, i0, i1, l0, l1);
    s.c_lo = i0;
    src_index(g.sx, s.xe - 1, g.w, i0, i1, l0, l1);
    s.nc = i1 - s.c_lo + 1;
    src_index(g.sy, s.ys, g.h, i0, i1, l0, l1);
    s.r_lo = i0;
    src_index(g.sy, s.ye - 1, g.h, i0, i1, l0, l1);
    s.nr = i1 - s.r_lo + 1;
    return s;
}

// stage the low-res tile [C][nr][nc] (pitches nrm, ncp) in shared memory
__device__ __forceinline__ void load_tile(float* s_tile, const float* __restrict__ lo, const FusedGeo& g,
                                          const Strip& s) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int rows = g.C * s.nr;
    const float* base = lo + ((long long)s.n * g.C * g.h + s.r_lo) * g.w + s.c_lo;
    for (int row = wid; row < rows; row += nw) {
        const int c = row / s.nr, r = row - c * s.nr;
        const float* src = base + ((long long)c * g.h + r) * g.w;
        float* dst = s_tile + (c * g.nrm + r) * g.ncp;
        for (int j = lane; j < s.nc; j += 32) dst[j] = __ldg(src + j);
    }
}

// t_r[c] = fma(A[c][r][x0], lx0, A[c][r][x1] * lx1)   (horizontal pass of ATen's formula)
template <int CT>
__device__ __forceinline__ void hline(float (&Hx)[CT], const float* s_tile, const FusedGeo& g, int rr, int j0,
                                      int j1, float lx0, float lx1) {
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        if (c < g.C) {
            const float* row = s_tile + (c * g.nrm + rr) * g.ncp;
            Hx[c] = __fmaf_rn(row[j0], lx0, __fmul_rn(row[j1], lx1));
        } else {
            Hx[c] = kPadLogit;
        }
    }
}

// ------------------------------------------------------------------ K1: forward
template <int CT, bool IW, bool HAS_LABEL>
__global__ void __launch_bounds__(kTW, 6)
fused_fwd_kernel(const float* __restrict__ lo, FusedGeo g, int n_img, const int64_t* __restrict__ label, float r32,
                 float omr32, int n_norm, State st) {
    extern __shared__ float s_tile[];
    __shared__ unsigned s_hist[MSQ_MAX_CLASSES];
    __shared__ unsigned long long s_sum[MSQ_MAX_CLASSES];
    __shared__ unsigned s_flags;
    const int tid = threadIdx.x;
    if (tid < MSQ_MAX_CLASSES) { s_hist[tid] = 0u; s_sum[tid] = 0ull; }
    if (tid == 0) s_flags = 0u;
    const Strip sp = make_strip(g);
    load_tile(s_tile, lo, g, sp);
    __syncthreads();

    const bool active = (sp.xs + tid) < sp.xe;
    const int x = active ? sp.xs + tid : sp.xe - 1;
    int x0, x1;
    float lx0, lx1;
    src_index(g.sx, x, g.w, x0, x1, lx0, lx1);
    const int j0 = x0 - sp.c_lo, j1 = x1 - sp.c_lo;

    float Ha[CT], Hb[CT];
    int ra = -1, rb = -1;
    int run_k = -1;
    unsigned run_cnt = 0u;
    float run_q = 0.f;
    auto flush = [&]() {
        if (run_k >= 0 && run_cnt) {
            if (IW && !HAS_LABEL) atomicAdd(&s_hist[run_k], run_cnt);
            if (!(fabsf(run_q) < 3.0e38f)) atomicOr(&s_flags, kFlagNonFinite);
            if (IW) atomicAdd(&s_sum[run_k], to_fix(run_q));
        }
    };

    for (int y = sp.ys; y < sp.ye; ++y) {
        int y0, y1;
        float ly0, ly1;
        src_index(g.sy, y, g.h, y0, y1, ly0, ly1);
        if (y0 != ra) {
            if (y0 == rb) {
#pragma unroll
                for (int c = 0; c < CT; ++c) Ha[c] = Hb[c];
            } else {
                hline<CT>(Ha, s_tile, g, y0 - sp.r_lo, j0, j1, lx0, lx1);
            }
            ra = y0;
        }
        if (y1 != rb) {
            if (y1 == ra) {
#pragma unroll
                for (int c = 0; c < CT; ++c) Hb[c] = Ha[c];
            } else {
                hline<CT>(Hb, s_tile, g, y1 - sp.r_lo, j0, j1, lx0, lx1);
            }
            rb = y1;
        }
        float z[CT], e[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) z[c] = __fmaf_rn(Ha[c], ly0, __fmul_rn(Hb[c], ly1));
        float inv_s, q;
        const int k = pixel_stats<CT, IW>(z, e, inv_s, q);
        if (active) {
            if (IW) {
                if (HAS_LABEL) {
                    const long long lv = label[((long long)sp.n * g.H + y) * g.W + x];
                    if (lv >= 0 && lv < g.C) atomicAdd(&s_hist[(int)lv], 1u);
                }
                if (k == run_k) { run_cnt++; run_q += q; }
                else { flush(); run_k = k; run_cnt = 1u; run_q = q; }
            } else {
                run_k = 0; run_cnt++; run_q += q;
            }
        }
    }
    if (IW) {
        flush();
    } else {
        // MaxSquare: one bucket; reduce the strip in registers first
        if (!(fabsf(run_q) < 3.0e38f)) atomicOr(&s_flags, kFlagNonFinite);
        unsigned long long v = warp_sum_u64(run_cnt ? to_fix(run_q) : 0ull);
        if ((tid & 31) == 0 && v) atomicAdd(&s_sum[0], v);
    }
    __syncthreads();
    if (tid < g.C) {
        if (s_hist[tid]) atomicAdd(&st.hist[sp.n * g.C + tid], s_hist[tid]);
        if (s_sum[tid]) atomicAdd(&st.sumsq[sp.n * g.C + tid], s_sum[tid]);
    }
    if (tid == 0 && s_flags) atomicOr(st.flags, s_flags);
    if (take_ticket_is_last(st.ticket, gridDim.x * gridDim.y * gridDim.z))
        finalize_loss(st, IW ? MSQ_MODE_IW : MSQ_MODE_MAXSQUARE, n_img, g.C, r32, omr32, n_norm,
                      (unsigned long long)n_img * g.C * g.H * g.W);
}

// ------------------------------------------------------------------ K2: backward
// dL/dz_c = a * p_c * (p_c - q),  a = -2 w[n,k] go / (Nn C)  (IW)   or   -go / (Nn C H W)  (MaxSquare)
template <int CT, bool IW>
__global__ void __launch_bounds__(kTW, 4)
fused_bwd_kernel(const float* __restrict__ lo, FusedGeo g, int n_img, int n_norm, const float* __restrict__ weights,
                 const float* __restrict__ grad_out, float* __restrict__ grad_lo) {
    extern __shared__ float s_dyn[];
    float* s_tile = s_dyn;                                   // [C][nrm][ncp]
    float* s_stage = s_tile + g.C * g.nrm * g.ncp;           // [C][kTW+1]
    float* s_lx0 = s_stage + g.C * (kTW + 1);                // [kTW]
    float* s_lx1 = s_lx0 + kTW;                              // [kTW]
    int* s_j0 = (int*)(s_lx1 + kTW);                         // [kTW]
    int* s_j1 = s_j0 + kTW;                                  // [kTW]
    int* s_rng = s_j1 + kTW;                                 // [4][ncp]: start0,end0,start1,end1
    __shared__ float s_coef[MSQ_MAX_CLASSES];
    const int tid = threadIdx.x;
    const Strip sp = make_strip(g);
    const float go = *grad_out;
    if (IW && tid < g.C)
        s_coef[tid] = (float)(-2.0 * (double)weights[sp.n * g.C + tid] * (double)go / ((double)n_norm * (double)g.C));
    const float coef_ms = (float)(-(double)go / ((double)n_norm * (double)g.C * (double)g.H * (double)g.W));
    load_tile(s_tile, lo, g, sp);

    const bool active = (sp.xs + tid) < sp.xe;
    const int x = active ? sp.xs + tid : sp.xe - 1;
    int x0, x1;
    float lx0, lx1;
    src_index(g.sx, x, g.w, x0, x1, lx0, lx1);
    const int j0 = x0 - sp.c_lo, j1 = x1 - sp.c_lo;
    s_lx0[tid] = lx0;
    s_lx1[tid] = lx1;
    s_j0[tid] = active ? j0 : -1;
    s_j1[tid] = active ? j1 : -1;
    for (int i = tid; i < 4 * g.ncp; i += kTW) s_rng[i] = 0;
    __syncthreads();
    if (active) {
        const bool last = (sp.xs + tid + 1 == sp.xe);
        if (tid == 0 || s_j0[tid - 1] != j0) s_rng[0 * g.ncp + j0] = tid;
        if (last || s_j0[tid + 1] != j0) s_rng[1 * g.ncp + j0] = tid + 1;
        if (tid == 0 || s_j1[tid - 1] != j1) s_rng[2 * g.ncp + j1] = tid;
        if (last || s_j1[tid + 1] != j1) s_rng[3 * g.ncp + j1] = tid + 1;
    }
    __syncthreads();

    float Ha[CT], Hb[CT], dHa[CT], dHb[CT];
    int ra = -1, rb = -1;

    // horizontal adjoint of one finished low-res row: every thread parks its d t_r[c] in
    // shared memory, then one thread per (class, low-res column) gathers its <=2 runs
    // of output columns and issues one red.global.add.
    auto flush_row = [&](int r, const float (&dH)[CT]) {
#pragma unroll
        for (int c = 0; c < CT; ++c)
            if (c < g.C) s_stage[c * (kTW + 1) + tid] = active ? dH[c] : 0.f;
        __syncthreads();
        float* out = grad_lo + (((long long)sp.n * g.C) * g.h + r) * g.w + sp.c_lo;
        const int Cd = (CT == 13 || CT == 16 || CT == 19) ? CT : g.C;   // exact instantiations: constant divisor
        const int cells = Cd * sp.nc;
        for (int idx = tid; idx < cells; idx += kTW) {
            const int j = idx / Cd, c = idx - j * Cd;
            const float* col = s_stage + c * (kTW + 1);
            float acc = 0.f;
            for (int t = s_rng[j], te = s_rng[g.ncp + j]; t < te; ++t) acc = fmaf(s_lx0[t], col[t], acc);
            for (int t = s_rng[2 * g.ncp + j], te = s_rng[3 * g.ncp + j]; t < te; ++t) acc = fmaf(s_lx1[t], col[t], acc);
            atomicAdd(out + (long long)c * g.h * g.w + j, acc);
        }
        __syncthreads();
    };

    for (int y = sp.ys; y < sp.ye; ++y) {
        int y0, y1;
        float ly0, ly1;
        src_index(g.sy, y, g.h, y0, y1, ly0, ly1);
        if (y0 != ra) {
            if (ra >= 0) flush_row(ra, dHa);
            if (y0 == rb) {
#pragma unroll
                for (int c = 0; c < CT; ++c) { Ha[c] = Hb[c]; dHa[c] = dHb[c]; }
            } else {
                if (rb >= 0) flush_row(rb, dHb);
                hline<CT>(Ha, s_tile, g, y0 - sp.r_lo, j0, j1, lx0, lx1);
#pragma unroll
                for (int c = 0; c < CT; ++c) dHa[c] = 0.f;
            }
            ra = y0;
            rb = -1;
        }
        if (y1 != rb) {
            if (rb >= 0) flush_row(rb, dHb);
            if (y1 == ra) {
#pragma unroll
                for (int c = 0; c < CT; ++c) Hb[c] = Ha[c];
            } else {
                hline<CT>(Hb, s_tile, g, y1 - sp.r_lo, j0, j1, lx0, lx1);
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) dHb[c] = 0.f;
            rb = y1;
        }
        float z[CT], e[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) z[c] = __fmaf_rn(Ha[c], ly0, __fmul_rn(Hb[c], ly1));
        float inv_s, q;
        const int k = pixel_stats<CT, IW>(z, e, inv_s, q);
        const float a = IW ? s_coef[k] : coef_ms;
        const float a0 = a * ly0, a1 = a * ly1;
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            const float p = e[c] * inv_s;
            const float gz = p * (p - q);
            dHa[c] = fmaf(a0, gz, dHa[c]);
            dHb[c] = fmaf(a1, gz, dHb[c]);
        }
    }
    if (ra >= 0) flush_row(ra, dHa);
    if (rb >= 0) flush_row(rb, dHb);
}

// ------------------------------------------------------------------ host side
int g_fused_rows = 0;      // tuning knob: 0 = choose automatically

static int make_geo(int C, int h, int w, int H, int W, int n, FusedGeo& g) {
    if (H < h || W < w) return MSQ_E_GEOMETRY;
    g.C = C; g.h = h; g.w = w; g.H = H; g.W = W;
    g.sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    g.sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const int tiles_x = (W + kTW - 1) / kTW;
    int R = g_fused_rows;
    if (R <= 0) {
        // ~6 CTAs per SM in flight, strips no shorter than 4 rows and no longer than 32
        const long long target = 6LL * kSMs;
        R = (int)(((long long)H * tiles_x * n) / target);
        if (R < 4) R = 4;
        if (R > 32) R = 32;
    }
    if (R > H) R = H;
    g.R = R;
    // exact tile extents with the kernel's own index arithmetic
    int nrm = 1, ncp = 1, i0, i1, a0;
    float l0, l1;
    for (int ys = 0; ys < H; ys += R) {
        const int ye = (ys + R < H) ? ys + R : H;
        src_index(g.sy, ys, h, a0, i1, l0, l1);
        src_index(g.sy, ye - 1, h, i0, i1, l0, l1);
        if (i1 - a0 + 1 > nrm) nrm = i1 - a0 + 1;
    }
    for (int xs = 0; xs < W; xs += kTW) {
        const int xe = (xs + kTW < W) ? xs + kTW : W;
        src_index(g.sx, xs, w, a0, i1, l0, l1);
        src_index(g.sx, xe - 1, w, i0, i1, l0, l1);
        if (i1 - a0 + 1 > ncp) ncp = i1 - a0 + 1;
    }
    g.nrm = nrm;
    g.ncp = ncp | 1;       // odd pitch: consecutive tile rows start in different banks
    return 0;
}

static dim3 fused_grid(const FusedGeo& g, int n) {
    return dim3((unsigned)((g.W + kTW - 1) / kTW), (unsigned)((g.H + g.R - 1) / g.R), (unsigned)n);
}

template <int CT>
static int launch_fused_fwd(int mode, const float* lo, const FusedGeo& g, int n, const int64_t* label, float r32,
                            float omr32, int nn, State st, cudaStream_t s) {
    const size_t smem = (size_t)g.C * g.nrm * g.ncp * sizeof(float);
    if (smem > 200 * 1024) return MSQ_E_SMEM;
    const dim3 grid = fused_grid(g, n);
#define MSQ_LAUNCH(K)                                                                             \
    do {                                                                                          \
        if (smem > 48 * 1024) cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        K<<<grid, kTW, smem, s>>>(lo, g, n, label, r32, omr32, nn, st);                           \
    } while (0)
    if (mode == MSQ_MODE_MAXSQUARE) MSQ_LAUNCH((fused_fwd_kernel<CT, false, false>));
    else if (label) MSQ_LAUNCH((fused_fwd_kernel<CT, true, true>));
    else MSQ_LAUNCH((fused_fwd_kernel<CT, true, false>));
#undef MSQ_LAUNCH
    MSQ_CHECK_LAUNCH();
    return 0;
}

template <int CT>
static int launch_fused_bwd(int mode, const float* lo, const FusedGeo& g, int n, int nn, State st,
                            const float* grad_out, float* grad_lo, cudaStream_t s) {
    const size_t smem = ((size_t)g.C * g.nrm * g.ncp + (size_t)g.C * (kTW + 1) + 2 * kTW) * sizeof(float) +
                        (2 * kTW + 4 * (size_t)g.ncp) * sizeof(int);
    if (smem > 200 * 1024) return MSQ_E_SMEM;
    const dim3 grid = fused_grid(g, n);
    cudaError_t e = cudaMemsetAsync(grad_lo, 0, (size_t)n * g.C * g.h * g.w * sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
#define MSQ_LAUNCH(K)                                                                             \
    do {                                                                                          \
        if (smem > 48 * 1024) cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        K<<<grid, kTW, smem, s>>>(lo, g, n, nn, st.weights, grad_out, grad_lo);                   \
    } while (0)
    if (mode == MSQ_MODE_MAXSQUARE) MSQ_LAUNCH((fused_bwd_kernel<CT, false>));
    else MSQ_LAUNCH((fused_bwd_kernel<CT, true>));
#undef MSQ_LAUNCH
    MSQ_CHECK_LAUNCH();
    return 0;
}

}  // namespace msq

using namespace msq;

#define MSQ_DISPATCH_C(C, CALL)                  \
    switch (C) {                                 \
        case 13: return CALL(13);                \
        case 16: return CALL(16);                \
        case 19: return CALL(19);                \
        default:                                 \
            if ((C) <= 8) return CALL(8);        \
            if ((C) <= 24) return CALL(24);      \
            return CALL(32);                     \
    }

extern "C" int msq_fused_fwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                             const int64_t* label, double ratio, int n_images_norm, void* accum, void* out,
                             msq_stream_t stream) {
    if (!logits || !accum || !out || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES || h < 1 || w < 1 || out_h < 1 ||
        out_w < 1)
        return MSQ_E_BADARG;
    if (mode != MSQ_MODE_IW && mode != MSQ_MODE_MAXSQUARE) return MSQ_E_BADARG;
    if ((((uintptr_t)logits) & 3u) || ((((uintptr_t)accum) | ((uintptr_t)out)) & 15u) || (label && (((uintptr_t)label) & 7u))) return MSQ_E_ALIGN;
    FusedGeo g;
    const int rc = make_geo(num_class, h, w, out_h, out_w, n, g);
    if (rc) return rc;
    const State st = carve(accum, out, n, num_class);
    const float r32 = (float)ratio, omr32 = (float)(1.0 - ratio);
    const int nn = n_images_norm > 0 ? n_images_norm : n;
    cudaStream_t s = (cudaStream_t)stream;
#define CALL(CT) launch_fused_fwd<CT>(mode, logits, g, n, label, r32, omr32, nn, st, s)
    MSQ_DISPATCH_C(num_class, CALL)
#undef CALL
}

extern "C" int msq_fused_bwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                             int n_images_norm, const void* out, const float* grad_out, float* grad_logits,
                             msq_stream_t stream) {
    if (!logits || !out || !grad_out || !grad_logits || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES ||
        h < 1 || w < 1 || out_h < 1 || out_w < 1)
        return MSQ_E_BADARG;
    if (mode != MSQ_MODE_IW && mode != MSQ_MODE_MAXSQUARE) return MSQ_E_BADARG;
    if ((((uintptr_t)logits) | ((uintptr_t)grad_logits) | ((uintptr_t)grad_out)) & 3u) return MSQ_E_ALIGN;
    FusedGeo g;
    const int rc = make_geo(num_class, h, w, out_h, out_w, n, g);
    if (rc) return rc;
    const State st = carve(nullptr, const_cast<void*>(out), n, num_class);
    const int nn = n_images_norm > 0 ? n_images_norm : n;
    cudaStream_t s = (cudaStream_t)stream;
#define CALL(CT) launch_fused_bwd<CT>(mode, logits, g, n, nn, st, grad_out, grad_logits, s)
    MSQ_DISPATCH_C(num_class, CALL)
#undef CALL
}
