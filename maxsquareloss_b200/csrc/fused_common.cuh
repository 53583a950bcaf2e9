// Shared device code of the fused kernels (fused_loss.cu, guidance.cu): geometry, tile
// staging, per-pixel softmax statistics, the forward / backward column-walk kernels and the
// host-side launch planning.
#pragma once
#include <mutex>
#include "common.cuh"


namespace msq {

#ifndef MSQ_TW
#define MSQ_TW 128                              // measured: 64-thread CTAs (8 per SM) lose at batch 1-2 (40.4 vs 33.9 us), profiles/r02_ab_rejected.txt
#endif
constexpr int kTW = MSQ_TW;                     // output columns (= threads) per CTA
static_assert(kTW % 32 == 0 && kTW >= 32 && kTW <= 256, "kTW: whole warps, at least MSQ_MAX_CLASSES threads");
constexpr int kRun = 8;                         // backward fast path: output columns per low-res cell and tap side
#ifndef MSQ_FWD_MINB
#define MSQ_FWD_MINB 4                          // co-resident CTAs per SM the forward is compiled for
#endif
#ifndef MSQ_BWD_MINB
#define MSQ_BWD_MINB 4
#endif
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
    unsigned uq, ur;     // units = uq * grid + ur: CTA b owns [b*uq + b*ur/grid, (b+1)*uq + (b+1)*ur/grid)  (b*ur < 2^32)
    unsigned zq, zr;     // the same split of the n*C*h*w elements of dL/dlogits that the forward zero-fills
    int fastx;           // every run of output columns sharing x0 (or x1) within a tile is <= kRun long
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

// Classes are processed two at a time with Blackwell's packed fp32x2 instructions
// (fma.rn.f32x2 / mul.f32x2 / add.f32x2, sm_100+): the kernels are instruction-issue
// bound, and one FFMA2 does the work of two FFMAs with identical IEEE rounding per
// lane, so bit-exactness of the interpolated logits is unaffected.  Class c lives in
// lane (c & 1) of pair (c >> 1); an odd class count leaves one padding lane.
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
__device__ __forceinline__ float lane_of(const float2& v, int c) { return (c & 1) ? v.y : v.x; }
// 2^t for the two classes of pair p.  With an odd class count the last pair's second lane is padding (logit -1e30,
// exponential 0): no MUFU is spent on it (the kernels are MUFU co-limited: 20 -> 19 ex2 per pixel at C = 19).
template <int CT>
__device__ __forceinline__ float2 ex2_pair(const float2& t, int p) {
    return make_float2(ex2_approx(t.x), (2 * p + 1 < CT) ? ex2_approx(t.y) : 0.f);
}

// Near-maximum class word of a pixel: bit CT-1-c is set when class c is within 2^-22 of the maximum m.
//   z_c >= thr  <=>  sign(z_c - thr) == 0   (round-to-nearest: the sign of a difference is exact, a zero difference is +0)
// so one packed FADD2 per class pair forms the differences and one funnel shift per class appends the sign bit to a
// word: 10 + 19 instructions where a compare + predicated OR per class took 38 (measured: forward 17.6 -> 17.2 us).
template <int CT>
__device__ __forceinline__ unsigned pixel_near(const float2 (&z)[(CT + 1) / 2], float thr) {
    constexpr int CP = (CT + 1) / 2;
    const float2 nthr = splat(-thr);
    unsigned acc = 0u;
#pragma unroll
    for (int p = 0; p < CP; ++p) {
        const float2 d = __fadd2_rn(z[p], nthr);
        acc = __funnelshift_l(__float_as_uint(d.x), acc, 1);
        if (2 * p + 1 < CT) acc = __funnelshift_l(__float_as_uint(d.y), acc, 1);
    }
    return ~acc & (CT >= 32 ? 0xffffffffu : ((1u << (CT & 31)) - 1u));
}
// the FIRST class of the word (its highest bit); several candidates: replay torch's arithmetic (resolve_ties)
template <int CT>
__device__ __forceinline__ int near_to_class(unsigned near, const float2 (&z)[(CT + 1) / 2], float m) {
    int k = CT - 1 - (31 - __clz((int)near));
    if (near & (near - 1u)) {
        asm volatile("" ::: "memory");                  // keep the spill of z[] inside this cold branch
        float zl[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) zl[c] = lane_of(z[c >> 1], c);
        k = resolve_ties<CT>(zl, m);
    }
    return min(max(k, 0), CT - 1);                      // NaN logits: any valid class (the reference yields a NaN loss anyway)
}

// Per-pixel softmax statistics from the interpolated logits z[] (pairs):
//   e[c] = 2^((z_c - m) log2 e), inv_s = 1/s with s = sum e, q = sum_c p_c^2, qs = q*s;
//   returns the argmax class.
template <int CT, bool NEED_ARG>
__device__ __forceinline__ int pixel_stats(const float2 (&z)[(CT + 1) / 2], float2 (&e)[(CT + 1) / 2], float& m_out,
                                           float& inv_s, float& q, float& qs) {
    constexpr int CP = (CT + 1) / 2;
    float m = z[0].x;
#pragma unroll
    for (int c = 1; c < CT; ++c) m = fmaxf(m, lane_of(z[c >> 1], c));
    m_out = m;
    int k = 0;
    if (NEED_ARG) {
        k = near_to_class<CT>(pixel_near<CT>(z, m - kNearTie), z, m);
    }
    const float2 l2e = splat(kLog2e), nm = splat(-m * kLog2e);
    float2 s2a = make_float2(0.f, 0.f), s2b = s2a, ss2a = s2a, ss2b = s2a;
#pragma unroll
    for (int p = 0; p < CP; ++p) {
        const float2 t = __ffma2_rn(z[p], l2e, nm);
        e[p] = ex2_pair<CT>(t, p);
        if (p & 1) { s2b = __fadd2_rn(s2b, e[p]); ss2b = __ffma2_rn(e[p], e[p], ss2b); }
        else { s2a = __fadd2_rn(s2a, e[p]); ss2a = __ffma2_rn(e[p], e[p], ss2a); }
    }
    const float2 s2 = __fadd2_rn(s2a, s2b), ss2 = __fadd2_rn(ss2a, ss2b);
    const float s = s2.x + s2.y, ss = ss2.x + ss2.y;
    inv_s = rcp_approx(s);
    qs = ss * inv_s;            // q * s
    q = qs * inv_s;
    return k;
}

// MinEnt variant (utils/loss.py:17-67, softCrossEntropy / IWsoftCrossEntropy called with
// target = softmax(inputs)): per pixel the entropy H = -sum_c p_c log p_c = ln s - A/s with
// t_c = z_c - m, e_c = exp(t_c), s = sum e_c, A = sum e_c t_c (both terms are >= 0: no cancellation).
// Returns H in `q`, D = A/s in `qs` and 1/s in `inv_s`;  dH/dz_j = -p_j (t_j - D).
// The argmax is the reference's torch.max(inputs, 1) on the LOGITS (utils/loss.py:54): first maximum.
template <int CT, bool NEED_ARG>
__device__ __forceinline__ int pixel_stats_entropy(const float2 (&z)[(CT + 1) / 2], float& m_out, float& inv_s, float& q,
                                                   float& qs) {
    constexpr int CP = (CT + 1) / 2;
    float m = z[0].x;
#pragma unroll
    for (int c = 1; c < CT; ++c) m = fmaxf(m, lane_of(z[c >> 1], c));
    m_out = m;
    int k = 0;
    if (NEED_ARG) {
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
        k = __ffs(mask_a | mask_b) - 1;
        if (k < 0) k = 0;
    }
    const float2 l2e = splat(kLog2e), nm2 = splat(-m);
    float2 sa = make_float2(0.f, 0.f), sb = sa, aa = sa, ab = sa;
#pragma unroll
    for (int p = 0; p < CP; ++p) {
        const float2 t = __fadd2_rn(z[p], nm2);
        const float2 tl = __fmul2_rn(t, l2e);
        const float2 e = ex2_pair<CT>(tl, p);
        // padded lanes: t = -1e30, e = 0 exactly; keep 0 * -1e30 out of the sum
        const float2 tc = make_float2(fmaxf(t.x, -1.0e4f), fmaxf(t.y, -1.0e4f));
        if (p & 1) { sb = __fadd2_rn(sb, e); ab = __ffma2_rn(e, tc, ab); }
        else { sa = __fadd2_rn(sa, e); aa = __ffma2_rn(e, tc, aa); }
    }
    const float2 s2 = __fadd2_rn(sa, sb), a2 = __fadd2_rn(aa, ab);
    const float s = s2.x + s2.y, A = a2.x + a2.y;
    inv_s = rcp_approx(s);
    qs = A * inv_s;                 // D  (<= 0)
    q = logf(s) - qs;               // H  (>= 0)
    return k;
}

// Optional per-pixel statistics cache written by the forward and read by the backward: one
// float4 per pixel {max logit m, q*s, 1/s^2, argmax class (int bits)} = 16 B/pixel, one
// coalesced 128-bit store / load per thread and row.  With it the backward skips the max /
// near-tie mask / sum / reciprocal work (~55 % of its instructions) and needs no reduction
// over the classes at all; without it (aux == NULL) it recomputes everything from the logits.

// ---------------------------------------------------------------------------------
// Work partition.  The output is cut into "row units": one unit = one output row of
// one kTW-wide column tile of one image, numbered u = (n*TX + tx)*H + y.  The grid
// has exactly as many CTAs as fit on the chip at once (148 x occupancy, one wave, no
// tail) and CTA b owns the contiguous unit range [b*U/G, (b+1)*U/G): every CTA gets
// the same number of rows to within one.  A range is walked as 1..2 "segments"
// (it may cross a column-tile boundary); per segment the low-res tile is staged in
// shared memory and each thread walks its column down the rows.
// ---------------------------------------------------------------------------------
struct Strip {
    int n, xs, xe, ys, ye;        // output extent of this segment
    int c_lo, r_lo, nc, nr;       // low-res tile origin / extent
};

__device__ __forceinline__ Strip make_strip(const FusedGeo& g, int col, int TX, int ys, int ye) {
    Strip s;
    s.n = col / TX;
    s.xs = (col - s.n * TX) * kTW;
    s.xe = min(g.W, s.xs + kTW);
    s.ys = ys;
    s.ye = ye;
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
    if (s.nc > g.ncp || s.nr > g.nrm) __trap();      // host sized the tile from the same arithmetic
    return s;
}

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Class pitch of the shared-memory tile: the tile is stored channel-LAST, [row][col][cpd(CT)],
// so that the C logits of one low-res cell are contiguous and a thread fetches them with
// cpd/4 LDS.128 instead of C scalar loads with C address computations.
__host__ __device__ constexpr int cpd(int ct) { return (ct + 3) / 4 * 4; }

// Stage the low-res tile [nr][nc][CPD] (row pitch ncp cells) in shared memory with
// cp.async (LDGSTS): every element is in flight at once, no register staging.
// Thread t owns tile cell (t / nc, t % nc): one integer division per segment, then one
// cp.async per class (global reads coalesced along the row across threads).  Lanes c >= C of
// the cell (padded class counts, and the odd lane of an odd C) get kPadLogit from the same thread.
template <int CT, bool PAD>
__device__ __forceinline__ void load_tile_issue(float* s_tile, const float* __restrict__ lo, const FusedGeo& g,
                                                const Strip& s) {
    constexpr int CPD = cpd(CT);
    const int cells = s.nr * s.nc;
    const int hw = g.h * g.w;
    const float* base = lo + ((long long)s.n * g.C * g.h + s.r_lo) * g.w + s.c_lo;
    for (int t = threadIdx.x; t < cells; t += kTW) {
        const int r = t / s.nc, j = t - r * s.nc;
        const float* src = base + r * g.w + j;
        float* dst = s_tile + (r * g.ncp + j) * CPD;
        const int Cn = PAD ? g.C : CT;
#pragma unroll 4
        for (int c = 0; c < Cn; ++c) cp_async4(dst + c, src + (long long)c * hw);
#pragma unroll
        for (int c = CT; c < CPD; ++c) dst[c] = kPadLogit;
        if (PAD) for (int c = g.C; c < CT; ++c) dst[c] = kPadLogit;
    }
    cp_async_commit();
}
template <int CT, bool PAD>
__device__ __forceinline__ void load_tile(float* s_tile, const float* __restrict__ lo, const FusedGeo& g, const Strip& s) {
    load_tile_issue<CT, PAD>(s_tile, lo, g, s);
    cp_async_wait<0>();
}

// t_r[c] = fma(A[c][r][x0], lx0, A[c][r][x1] * lx1)   (horizontal pass of ATen's formula)
template <int CT, bool PAD>
__device__ __forceinline__ void hline(float2 (&Hx)[(CT + 1) / 2], const float* s_tile, const FusedGeo& g, int rr,
                                      int j0, int j1, float lx0, float lx1) {
    constexpr int CP = (CT + 1) / 2, CPD = cpd(CT);
    const float4* c0 = reinterpret_cast<const float4*>(s_tile + (rr * g.ncp + j0) * CPD);
    const float4* c1 = reinterpret_cast<const float4*>(s_tile + (rr * g.ncp + j1) * CPD);
    const float2 l0 = splat(lx0), l1 = splat(lx1);
#pragma unroll
    for (int q = 0; q < CPD / 4; ++q) {
        const float4 a = c0[q], b = c1[q];
        Hx[2 * q] = __ffma2_rn(make_float2(a.x, a.y), l0, __fmul2_rn(make_float2(b.x, b.y), l1));
        if (2 * q + 1 < CP) Hx[2 * q + 1] = __ffma2_rn(make_float2(a.z, a.w), l0, __fmul2_rn(make_float2(b.z, b.w), l1));
    }
}

// Per-segment table of the vertical interpolation parameters (identical for every thread of
// the CTA): row y -> {ly0, ly1, y0, y1}.  One broadcast LDS.128 per row instead of the
// I2F / FMUL / F2I / clamp sequence (3 XU-pipe conversions) in every thread.
constexpr int kRowTabMax = 512;
__device__ __forceinline__ void fill_row_table(float4* s_rows, const FusedGeo& g, int ys, int ye) {
    for (int r = threadIdx.x; r < ye - ys; r += kTW) {
        int y0, y1;
        float ly0, ly1;
        src_index(g.sy, ys + r, g.h, y0, y1, ly0, ly1);
        s_rows[r] = make_float4(ly0, ly1, __int_as_float(y0), __int_as_float(y1));
    }
}
__device__ __forceinline__ void row_params(const float4* s_rows, const FusedGeo& g, bool use_tab, int ys, int y,
                                           int& y0, int& y1, float& ly0, float& ly1) {
    if (use_tab) {
        const float4 v = s_rows[y - ys];
        ly0 = v.x; ly1 = v.y; y0 = __float_as_int(v.z); y1 = __float_as_int(v.w);
    } else {
        src_index(g.sy, y, g.h, y0, y1, ly0, ly1);
    }
}

// CTA-level accumulators of the image-wise statistics: per class a pixel count and the 64-bit
// fixed-point sum of q kept as two 32-bit words in shared memory.  Native 32-bit shared atomics
// (ATOMS.ADD); the low word's wrap-around is detected from the value the atomic returns and
// carried into the high word, so the pair is an exact 64-bit integer sum whatever the order.
// (Round 1 kept a private packed bucket per thread and class: 19 KB of shared memory per CTA whose
// zero-fill and shuffle reduction were 9 % of the forward's instructions and 21 % of its warp time
// at 14 rows per CTA -- profiles/r02_src_regions.txt.)
struct ClassAcc {
    unsigned cnt[MSQ_MAX_CLASSES], lo[MSQ_MAX_CLASSES], hi[MSQ_MAX_CLASSES];
};
__device__ __forceinline__ void class_acc_zero(ClassAcc& a, int tid) {
    if (tid < MSQ_MAX_CLASSES) { a.cnt[tid] = 0u; a.lo[tid] = 0u; a.hi[tid] = 0u; }
}
__device__ __forceinline__ void class_acc_add(ClassAcc& a, int k, unsigned long long fx, unsigned cnt, bool with_cnt) {
    const unsigned lo = (unsigned)fx;
    const unsigned old = atomicAdd(&a.lo[k], lo);
    const unsigned hi = (unsigned)(fx >> 32) + ((old + lo < old) ? 1u : 0u);
    if (hi) atomicAdd(&a.hi[k], hi);
    if (with_cnt) atomicAdd(&a.cnt[k], cnt);
}
// called by thread c < C after a __syncthreads(): fetch and clear class c
__device__ __forceinline__ void class_acc_take(ClassAcc& a, int c, unsigned& cnt, unsigned long long& sum) {
    cnt = a.cnt[c];
    sum = ((unsigned long long)a.hi[c] << 32) | (unsigned long long)a.lo[c];
    a.cnt[c] = 0u; a.lo[c] = 0u; a.hi[c] = 0u;
}

// ------------------------------------------------------------------ K1: forward
// IW: a thread follows the argmax class down its column and adds a finished run (pixel count, sum of q
//     in 2^-32 fixed point) to the CTA's per-class accumulators (ClassAcc, shared-memory atomics); when
//     the CTA leaves an image, thread c merges class c into the global replica with one atomic each.
template <int CT, bool PAD, bool IW, bool HAS_LABEL, int LOSS = 0>
__global__ void __launch_bounds__(kTW, MSQ_FWD_MINB)
fused_fwd_kernel(const float* __restrict__ lo, FusedGeo g, int n_img, unsigned units, const int64_t* __restrict__ label,
                 State st, void* __restrict__ aux, float* __restrict__ zero_buf, unsigned zero_count) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const bool use_tab = g.R <= kRowTabMax;
    float4* s_rows = (float4*)s_raw;                                          // [min(R, kRowTabMax)]
    float* s_tile = (float*)(s_rows + (use_tab ? g.R : 0));                   // [nrm][ncp][cpd(CT)]
    __shared__ ClassAcc s_acc;                                                // IW statistics of the current image
    __shared__ unsigned s_lab[MSQ_MAX_CLASSES];                               // label= histogram
    const int tid = threadIdx.x, lane = tid & 31;
    const int rep_off = (int)(blockIdx.x % kRep) * n_img * g.C;               // this CTA's accumulator replica
    MSQ_TRACE_PT((units >> 31), 0);
    pdl_trigger();          // the next kernel (backward or finalisation) may be scheduled as soon as SMs free up
    if (IW) class_acc_zero(s_acc, tid);
    if (HAS_LABEL && tid < MSQ_MAX_CLASSES) s_lab[tid] = 0u;
    float4* __restrict__ ax = (float4*)aux;

    // units < 2^31 (checked on the host): 32-bit divisions only
    unsigned u = blockIdx.x * g.uq + blockIdx.x * g.ur / gridDim.x;   // = floor(b * units / grid) without a 64-bit division
    const unsigned u_end = (blockIdx.x + 1) * g.uq + (blockIdx.x + 1) * g.ur / gridDim.x;
    const unsigned TX = (unsigned)((g.W + kTW - 1) / kTW);
    unsigned long long ms_acc = 0ull;
    bool bad = false, first = true;
    while (u < u_end) {
        const unsigned col = u / (unsigned)g.H;
        const int ys = (int)(u - col * (unsigned)g.H);
        const int ye = (int)min((unsigned)g.H, (unsigned)ys + (u_end - u));
        u += (unsigned)(ye - ys);
        const Strip sp = make_strip(g, (int)col, (int)TX, ys, ye);
        const bool active = (sp.xs + tid) < sp.xe;
        const int x = active ? sp.xs + tid : sp.xe - 1;
        int x0, x1;
        float lx0, lx1;
        src_index(g.sx, x, g.w, x0, x1, lx0, lx1);
        const int j0 = x0 - sp.c_lo, j1 = x1 - sp.c_lo;
        // launched with programmatic stream serialisation: the index arithmetic above overlaps the tail of whatever
        // precedes this kernel in the stream; global memory is touched only from here on
        if (first) {
            MSQ_TRACE_PT((units >> 31), 1);
            pdl_wait();
            MSQ_TRACE_PT((units >> 31), 2);
        }
        __syncthreads();                                   // previous segment done with s_tile / accumulators zeroed
        load_tile_issue<CT, PAD>(s_tile, lo, g, sp);       // cp.async in flight while the row table and the zero-fill are written
        if (use_tab) fill_row_table(s_rows, g, sp.ys, sp.ye);
        if (first && zero_buf) {                           // zero dL/dlogits for the backward's red.adds: no memset launch
            const unsigned z0 = blockIdx.x * g.zq + blockIdx.x * g.zr / gridDim.x;
            const unsigned z1 = (blockIdx.x + 1) * g.zq + (blockIdx.x + 1) * g.zr / gridDim.x;
            for (unsigned i = z0 + tid; i < z1; i += kTW) zero_buf[i] = 0.f;
        }
        first = false;
        cp_async_wait<0>();
        __syncthreads();
        MSQ_TRACE_PT((units >> 31), 3);

        constexpr int CP = (CT + 1) / 2;
        float2 Ha[CP], Hb[CP];
        int ra = -1, rb = -1;
        float4* axp = aux ? ax + (((long long)sp.n * g.H + sp.ys) * g.W + x) : nullptr;
        int run_k = -1;
        unsigned run_cnt = 0u;
        float run_q = 0.f;
        auto flush = [&]() {
            if (run_cnt) {
                bad |= !(fabsf(run_q) < 3.0e38f);
                if (IW) class_acc_add(s_acc, run_k, to_fix(run_q), run_cnt, !HAS_LABEL);
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
            float2 z[CP], e[CP];
            {
                const float2 w0 = splat(ly0), w1 = splat(ly1);
#pragma unroll
                for (int p = 0; p < CP; ++p) z[p] = __ffma2_rn(Ha[p], w0, __fmul2_rn(Hb[p], w1));
            }
            float m, inv_s, q, qs;
            const int k = (LOSS == 0) ? pixel_stats<CT, IW>(z, e, m, inv_s, q, qs)
                                      : pixel_stats_entropy<CT, IW>(z, m, inv_s, q, qs);
            if (active) {
                if (aux) axp[(long long)(y - sp.ys) * g.W] = make_float4(m, qs, LOSS == 0 ? inv_s * inv_s : inv_s, __int_as_float(k));
                if (IW) {
                    if (HAS_LABEL) {
                        const long long lv = label[((long long)sp.n * g.H + y) * g.W + x];
                        if (lv >= 0 && lv < g.C) atomicAdd(&s_lab[(int)lv], 1u);
                    }
                    if (k == run_k) { run_cnt++; run_q += q; }
                    else { flush(); run_k = k; run_cnt = 1u; run_q = q; }
                } else {
                    run_k = 0; run_cnt++; run_q += q;
                }
            }
        }
        flush();
        MSQ_TRACE_PT((units >> 31), 4);

        const int next_n = (u < u_end) ? (int)((u / (unsigned)g.H) / TX) : -1;
        if (IW) {
            // merge the CTA's accumulators into the global replica when the CTA is done with this image
            if (next_n == sp.n) continue;
            __syncthreads();
            if (tid < g.C) {
                unsigned cnt;
                unsigned long long sum;
                class_acc_take(s_acc, tid, cnt, sum);
                if (HAS_LABEL) { cnt = s_lab[tid]; s_lab[tid] = 0u; }
                if (cnt) atomicAdd(&st.hist[rep_off + sp.n * g.C + tid], cnt);
                if (sum) atomicAdd(&st.sumsq[rep_off + sp.n * g.C + tid], sum);
            }
        } else {
            // MaxSquare: hand the running sum over when the next segment belongs to another image
            if (next_n != sp.n) {
                ms_acc = warp_sum_u64(ms_acc);
                if (lane == 0 && ms_acc) atomicAdd(&st.sumsq[rep_off + sp.n * g.C], ms_acc);
                ms_acc = 0ull;
            }
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(st.flags, kFlagNonFinite);
    MSQ_TRACE_PT((units >> 31), 5);
}

// One-call step: the finalisation as extra CTAs of the backward grid (blockIdx >= the work grid `gw`).
//   which == 0: finalize_body, then -- once every work CTA has taken what it needs from the accumulators (each bumps the
//               `ticket` word after it has derived its last image's weights) -- the self-clean;
//   which == 1: the sharded step's statistics exchange over NVLink peer memory (PeerBox, common.cuh).
// (Inlined: passed to a non-inlined function the argument structure is copied to local memory by every thread of every CTA.)
static __device__ __forceinline__ void fin_cta(const FinArgs& fin, int which, unsigned gw) {
    pdl_wait();                      // the forward is complete
    pdl_trigger();
    if (which == 1) {
        if (threadIdx.x < 32) box_exchange(fin.box, (int)threadIdx.x);
        return;
    }
    finalize_body(fin.st, fin.mode, fin.n, fin.C, fin.r32, fin.omr32, fin.n_norm, fin.kept_dense, 0, fin.loss_kind, true);
    __syncthreads();
    if (fin.box.keep)                // this step's own [loss | hist] into the communicator's ring: pushed by the NEXT step's exchange
        for (int k = threadIdx.x; k < fin.box.keep_count; k += blockDim.x) fin.box.keep[k] = fin.st.stats[k];
    if (threadIdx.x == 0) {
        unsigned v, spins = 0u;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(fin.st.ticket) : "memory");
            if (v == gw) break;
            __nanosleep(100);
            if (++spins > (1u << 27)) __trap();      // > 10 s: a work CTA of this very grid never got to its weights (cannot happen)
        }
        *fin.st.ticket = 0u;
    }
    __syncthreads();
    finalize_clean(fin.st, fin.n, fin.C, 0);
}

// ------------------------------------------------------------------ K2: backward
// dL/dz_c = a * p_c * (p_c - q),  a = -2 w[n,k] go / (Nn C)  (IW)   or   -go / (Nn C H W)  (MaxSquare)
// GUIDE (needs CACHED): instead of the loss gradient, the gradient of the multi-level guidance
// cross-entropy on this head (tools/solve_gta5.py:213, nn.CrossEntropyLoss(ignore_index=-1)):
//   dL/dz_c = (go / n_valid) * (p_c - [c == label_2])   for pixels with label_2 != -1
// with {m, 1/s, label_2} read from the float4 cache msq_multi_fwd wrote for head 2.
template <int CT, bool PAD, bool IW, bool CACHED, bool GUIDE = false, int LOSS = 0>
__global__ void __launch_bounds__(kTW, MSQ_BWD_MINB)
fused_bwd_kernel(const float* __restrict__ lo, FusedGeo g, int n_img, unsigned units, int n_norm,
                 const float* __restrict__ weights, const float* __restrict__ grad_out, float grad_out_value,
                 float* __restrict__ grad_lo, const void* __restrict__ aux,
                 const unsigned long long* __restrict__ nvalid, const FinArgs fin) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr int CP = (CT + 1) / 2, CPD = cpd(CT), SP = CPD + 2;     // SP: stage pitch (floats), 8 B aligned rows
    const bool use_tab = g.R <= kRowTabMax;
    float4* s_rows = (float4*)s_raw;                         // [min(R, kRowTabMax)]
    float* s_tile = (float*)(s_rows + (use_tab ? g.R : 0));  // [nrm][ncp][CPD]
    float* s_stage = s_tile + CPD * g.nrm * g.ncp;           // [kTW + kRun][SP]: d t_r[c] of every column, class-innermost
    float* s_wt = s_stage + (kTW + kRun) * SP;               // [ncp][2 kRun]: horizontal weights of each cell's taps
    float* s_lx0 = s_wt + g.ncp * 2 * kRun;                  // [kTW]
    float* s_lx1 = s_lx0 + kTW;                              // [kTW]
    int* s_j0 = (int*)(s_lx1 + kTW);                         // [kTW]
    int* s_j1 = s_j0 + kTW;                                  // [kTW]
    int* s_rng = s_j1 + kTW;                                 // [4][ncp]: start0,end0,start1,end1
    __shared__ float s_coef[MSQ_MAX_CLASSES];
    const int tid = threadIdx.x;
    const bool fastx = g.fastx != 0;
    MSQ_TRACE_PT(2u + (units >> 31), 0);
    // fin.extra > 0 (one-call step, msq_fused_fwd_bwd): this kernel directly follows the forward.  Its work CTAs derive the
    // image-wise weights from the forward's replicated class histogram themselves (same arithmetic as the finalisation), and
    // the finalisation runs in an extra CTA BESIDE them instead of in a kernel between the two: a launch, two grid-completion
    // hand-overs and the finalisation's own load / powf / fp64 latency (4.5 us together) leave the step's critical path.
    const unsigned gw = gridDim.x - (unsigned)fin.extra;           // work grid
    if (blockIdx.x >= gw) {
        fin_cta(fin, (int)(blockIdx.x - gw), gw);
        return;
    }
    const unsigned* __restrict__ hist = fin.extra ? fin.st.hist : nullptr;
    if (!hist) pdl_trigger();
    for (int i = tid; i < kRun * SP; i += kTW) s_stage[kTW * SP + i] = 0.f;      // rows past the tile: zero taps
    // Launched with programmatic stream serialisation: everything up to pdl_wait() below (index
    // math, staging the logits tile, the column tables) overlaps the preceding kernel (the forward's tail in the
    // one-call step, the finalisation kernel otherwise); the upstream gradient, the weights / class histogram,
    // the statistics cache and dL/dlogits are touched only after it.
    float go = 0.f, coef_ms = 0.f;
    bool dep_ready = false;
    const int Cd = PAD ? g.C : CT;                           // exact instantiations: constant divisor
    const float4* __restrict__ ax = (const float4*)aux;

    // units < 2^31 (checked on the host): 32-bit divisions only
    unsigned u = blockIdx.x * g.uq + blockIdx.x * g.ur / gw;   // = floor(b * units / grid) without a 64-bit division
    const unsigned u_end = (blockIdx.x + 1) * g.uq + (blockIdx.x + 1) * g.ur / gw;
    const unsigned TX = (unsigned)((g.W + kTW - 1) / kTW);
    int coef_img = -1;
    while (u < u_end) {
        const unsigned col = u / (unsigned)g.H;
        const int ys = (int)(u - col * (unsigned)g.H);
        const int ye = (int)min((unsigned)g.H, (unsigned)ys + (u_end - u));
        u += (unsigned)(ye - ys);
        const Strip sp = make_strip(g, (int)col, (int)TX, ys, ye);
        __syncthreads();                                     // previous segment done with all shared arrays
        load_tile_issue<CT, PAD>(s_tile, lo, g, sp);         // the logits are complete before the step's first kernel (msq_b200.h)

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
        if (fastx) for (int i = tid; i < 2 * kRun * g.ncp; i += kTW) s_wt[i] = 0.f;
        if (use_tab) fill_row_table(s_rows, g, sp.ys, sp.ye);
        __syncthreads();
        if (active) {
            const bool last = (sp.xs + tid + 1 == sp.xe);
            if (tid == 0 || s_j0[tid - 1] != j0) s_rng[0 * g.ncp + j0] = tid;
            if (last || s_j0[tid + 1] != j0) s_rng[1 * g.ncp + j0] = tid + 1;
            if (tid == 0 || s_j1[tid - 1] != j1) s_rng[2 * g.ncp + j1] = tid;
            if (last || s_j1[tid + 1] != j1) s_rng[3 * g.ncp + j1] = tid + 1;
        }
        __syncthreads();
        if (fastx && active) {        // tap k of cell j: column s_rng[j] + k with weight lx (runs are <= kRun long: host-checked)
            const int k0 = tid - s_rng[j0], k1 = tid - s_rng[2 * g.ncp + j1];
            if (k0 < kRun) s_wt[j0 * 2 * kRun + k0] = lx0;
            if (k1 < kRun) s_wt[j1 * 2 * kRun + kRun + k1] = lx1;
        }
        // everything above is geometry and the logits tile: it overlaps the forward's tail and the finalisation.  The
        // upstream gradient, the weights, the statistics cache and dL/dlogits are touched only from here on.
        if (!dep_ready) {
            MSQ_TRACE_PT(2u + (units >> 31), 1);
            pdl_wait();
            if (hist) pdl_trigger();
            MSQ_TRACE_PT(2u + (units >> 31), 2);
        }
        // cached statistics of the first row: in flight together with the loads of the weights below
        const float4* axp = CACHED ? ax + (((long long)sp.n * g.H + sp.ys) * g.W + x) : nullptr;
        float4 nx = make_float4(0.f, 0.f, 0.f, 0.f);
        if (CACHED) nx = __ldg(axp);
        if (!dep_ready) {
            go = grad_out ? *grad_out : grad_out_value;      // device scalar (autograd) or by value (host pipeline)
            if (GUIDE) coef_ms = (float)((double)go / (double)(*nvalid));      // mean over the valid pixels
            else coef_ms = (float)(-(double)go / ((double)n_norm * (double)g.C * (double)g.H * (double)g.W));
            dep_ready = true;
        }
        if (IW && sp.n != coef_img) {
            if (hist) {
                if (tid < 32) {                              // utils/loss.py:92-96 for image sp.n, as finalize_body does it
                    unsigned hc = 0u;
                    if (tid < g.C) {
                        const int nc = n_img * g.C, idx = sp.n * g.C + tid;
#pragma unroll
                        for (int r = 0; r < kRep; ++r) hc += __ldcg(&hist[r * nc + idx]);
                    }
                    const unsigned total = __reduce_add_sync(0xffffffffu, hc);
                    if (tid < g.C)
                        s_coef[tid] = (float)((LOSS == 0 ? -2.0 : -1.0) * (double)iw_weight((float)hc, (float)total, fin.r32, fin.omr32) *
                                              (double)go / ((double)n_norm * (double)g.C));
                }
            } else if (tid < g.C) {
                s_coef[tid] = (float)((LOSS == 0 ? -2.0 : -1.0) * (double)weights[sp.n * g.C + tid] * (double)go /
                                      ((double)n_norm * (double)g.C));
            }
            coef_img = sp.n;
        }
        // last segment of this CTA: it needs nothing more from the accumulators (the loads above have been consumed)
        if (hist && u >= u_end && tid == 0) atomicAdd(fin.st.ticket, 1u);
        cp_async_wait<0>();
        __syncthreads();                                     // tile, tap tables and coefficients are in place

        MSQ_TRACE_PT(2u + (units >> 31), 3);
        float2 Ha[CP], Hb[CP], dHa[CP], dHb[CP];
        int ra = -1, rb = -1;

        // horizontal adjoint of one finished low-res row: every thread parks its d t_r[c] in
        // shared memory (class-innermost), then one thread per (low-res column, class pair)
        // gathers the <= 2 kRun taps of its cell with packed FMAs and issues the red.global.adds.
        auto flush_row = [&](int r, const float2 (&dH)[CP]) {
            float2* mine = reinterpret_cast<float2*>(s_stage + tid * SP);
#pragma unroll
            for (int p = 0; p < CP; ++p) mine[p] = active ? dH[p] : make_float2(0.f, 0.f);
            __syncthreads();
            float* out = grad_lo + (((long long)sp.n * g.C) * g.h + r) * g.w + sp.c_lo;
            const long long hw = (long long)g.h * g.w;
            if (fastx) {
                const int items = sp.nc * CP;
                for (int idx = tid; idx < items; idx += kTW) {
                    const int j = idx / CP, p = idx - j * CP;
                    const float4* wv = reinterpret_cast<const float4*>(s_wt + j * 2 * kRun);
                    const float2* a = reinterpret_cast<const float2*>(s_stage + s_rng[j] * SP) + p;
                    const float2* b = reinterpret_cast<const float2*>(s_stage + s_rng[2 * g.ncp + j] * SP) + p;
                    float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
#pragma unroll
                    for (int q = 0; q < kRun / 4; ++q) {
                        const float4 w0 = wv[q], w1 = wv[kRun / 4 + q];
                        acc0 = __ffma2_rn(splat(w0.x), a[(4 * q + 0) * (SP / 2)], acc0);
                        acc1 = __ffma2_rn(splat(w1.x), b[(4 * q + 0) * (SP / 2)], acc1);
                        acc0 = __ffma2_rn(splat(w0.y), a[(4 * q + 1) * (SP / 2)], acc0);
                        acc1 = __ffma2_rn(splat(w1.y), b[(4 * q + 1) * (SP / 2)], acc1);
                        acc0 = __ffma2_rn(splat(w0.z), a[(4 * q + 2) * (SP / 2)], acc0);
                        acc1 = __ffma2_rn(splat(w1.z), b[(4 * q + 2) * (SP / 2)], acc1);
                        acc0 = __ffma2_rn(splat(w0.w), a[(4 * q + 3) * (SP / 2)], acc0);
                        acc1 = __ffma2_rn(splat(w1.w), b[(4 * q + 3) * (SP / 2)], acc1);
                    }
                    const float2 acc = __fadd2_rn(acc0, acc1);
                    const int c0 = 2 * p;
                    if (!PAD || c0 < g.C) atomicAdd(out + (long long)c0 * hw + j, acc.x);
                    if (c0 + 1 < CT && (!PAD || c0 + 1 < g.C)) atomicAdd(out + (long long)(c0 + 1) * hw + j, acc.y);
                }
            } else {
                const int cells = Cd * sp.nc;
                for (int idx = tid; idx < cells; idx += kTW) {
                    const int j = idx / Cd, c = idx - j * Cd;
                    const float* colp = s_stage + c;
                    float acc = 0.f;
                    for (int t = s_rng[j], te = s_rng[g.ncp + j]; t < te; ++t) acc = fmaf(s_lx0[t], colp[t * SP], acc);
                    for (int t = s_rng[2 * g.ncp + j], te = s_rng[3 * g.ncp + j]; t < te; ++t) acc = fmaf(s_lx1[t], colp[t * SP], acc);
                    atomicAdd(out + (long long)c * hw + j, acc);
                }
            }
            __syncthreads();
        };

        // cached statistics of the next row are fetched while the current row is computed
        for (int y = sp.ys; y < sp.ye; ++y) {
            int y0, y1;
            float ly0, ly1;
            row_params(s_rows, g, use_tab, sp.ys, y, y0, y1, ly0, ly1);
            const float c_m = nx.x, c_qs = nx.y, c_is2 = nx.z;
            const float nx_prev_z = nx.z;                    // GUIDE: label_2 bits
            const int c_k = __float_as_int(nx.w);
            if (CACHED && y + 1 < sp.ye) { axp += g.W; nx = __ldg(axp); }
            if (y0 != ra) {
                if (ra >= 0) flush_row(ra, dHa);
                if (y0 == rb) {
#pragma unroll
                    for (int p = 0; p < CP; ++p) { Ha[p] = Hb[p]; dHa[p] = dHb[p]; }
                } else {
                    if (rb >= 0) flush_row(rb, dHb);
                    hline<CT, PAD>(Ha, s_tile, g, y0 - sp.r_lo, j0, j1, lx0, lx1);
#pragma unroll
                    for (int p = 0; p < CP; ++p) dHa[p] = make_float2(0.f, 0.f);
                }
                ra = y0;
                rb = -1;
            }
            if (y1 != rb) {
                if (rb >= 0) flush_row(rb, dHb);
                if (y1 == ra) {
#pragma unroll
                    for (int p = 0; p < CP; ++p) Hb[p] = Ha[p];
                } else {
                    hline<CT, PAD>(Hb, s_tile, g, y1 - sp.r_lo, j0, j1, lx0, lx1);
                }
#pragma unroll
                for (int p = 0; p < CP; ++p) dHb[p] = make_float2(0.f, 0.f);
                rb = y1;
            }
            // g_c = a p_c (p_c - q) = (a / s^2) e_c (e_c - q s)
            const float2 w0 = splat(ly0), w1 = splat(ly1);
            if (CACHED && GUIDE) {
                // cache = {m, 1/s, label_2}: p_c = e_c / s, g_c = coef * (p_c - onehot(label_2))
                const int lab = __float_as_int(nx_prev_z);
                const float a = (lab >= 0) ? coef_ms : 0.f;
                const float labf = (float)lab;
                const float2 a0 = splat(a * ly0), a1 = splat(a * ly1), is = splat(c_qs);
                const float2 l2e = splat(kLog2e), nm = splat(-c_m * kLog2e);
#pragma unroll
                for (int p = 0; p < CP; ++p) {
                    const float2 zp = __ffma2_rn(Ha[p], w0, __fmul2_rn(Hb[p], w1));
                    const float2 t = __ffma2_rn(zp, l2e, nm);
                    const float2 ep = ex2_pair<CT>(t, p);
                    const float2 noh = make_float2(labf == (float)(2 * p) ? -1.f : 0.f, labf == (float)(2 * p + 1) ? -1.f : 0.f);
                    const float2 v = __ffma2_rn(ep, is, noh);
                    dHa[p] = __ffma2_rn(a0, v, dHa[p]);
                    dHb[p] = __ffma2_rn(a1, v, dHb[p]);
                }
            } else if (CACHED && LOSS == 1) {
                // entropy: cache = {m, D, 1/s, k};  g_c = a e_c (t_c - D),  a = coef / s,  t_c = z_c - m
                const float a = (IW ? s_coef[c_k] : coef_ms) * c_is2;
                const float2 a0 = splat(a * ly0), a1 = splat(a * ly1), nD = splat(-c_qs);
                const float2 l2e = splat(kLog2e), nm2 = splat(-c_m);
#pragma unroll
                for (int p = 0; p < CP; ++p) {
                    const float2 zp = __ffma2_rn(Ha[p], w0, __fmul2_rn(Hb[p], w1));
                    const float2 t = __fadd2_rn(zp, nm2);
                    const float2 tl = __fmul2_rn(t, l2e);
                    const float2 ep = ex2_pair<CT>(tl, p);
                    const float2 tc = make_float2(fmaxf(t.x, -1.0e4f), fmaxf(t.y, -1.0e4f));
                    const float2 v = __fmul2_rn(ep, __fadd2_rn(tc, nD));
                    dHa[p] = __ffma2_rn(a0, v, dHa[p]);
                    dHb[p] = __ffma2_rn(a1, v, dHb[p]);
                }
            } else if (CACHED) {
                // max, q*s, 1/s^2 and the argmax class come from the forward's cache: no reduction over
                // the classes is left, so every class pair streams straight into the accumulators
                const float a = (IW ? s_coef[c_k] : coef_ms) * c_is2;
                const float2 a0 = splat(a * ly0), a1 = splat(a * ly1), nqs = splat(-c_qs);
                const float2 l2e = splat(kLog2e), nm = splat(-c_m * kLog2e);
#pragma unroll
                for (int p = 0; p < CP; ++p) {
                    const float2 zp = __ffma2_rn(Ha[p], w0, __fmul2_rn(Hb[p], w1));
                    const float2 t = __ffma2_rn(zp, l2e, nm);
                    const float2 ep = ex2_pair<CT>(t, p);
                    const float2 v = __fmul2_rn(ep, __fadd2_rn(ep, nqs));
                    dHa[p] = __ffma2_rn(a0, v, dHa[p]);
                    dHb[p] = __ffma2_rn(a1, v, dHb[p]);
                }
            } else {
                float2 z[CP], e[CP];
#pragma unroll
                for (int p = 0; p < CP; ++p) z[p] = __ffma2_rn(Ha[p], w0, __fmul2_rn(Hb[p], w1));
                float m, inv_s, q, qs;
                const int k = pixel_stats<CT, IW>(z, e, m, inv_s, q, qs);
                const float a = (IW ? s_coef[k] : coef_ms) * inv_s * inv_s;
                const float2 a0 = splat(a * ly0), a1 = splat(a * ly1), nqs = splat(-qs);
#pragma unroll
                for (int p = 0; p < CP; ++p) {
                    const float2 v = __fmul2_rn(e[p], __fadd2_rn(e[p], nqs));
                    dHa[p] = __ffma2_rn(a0, v, dHa[p]);
                    dHb[p] = __ffma2_rn(a1, v, dHb[p]);
                }
            }
        }
        MSQ_TRACE_PT(2u + (units >> 31), 4);
        if (ra >= 0) flush_row(ra, dHa);
        if (rb >= 0) flush_row(rb, dHb);
        MSQ_TRACE_PT(2u + (units >> 31), 5);
    }
}

// ------------------------------------------------------------------ host side
extern int g_fused_rows;   // tuning knob (fused_loss.cu): 0 = automatic (one balanced wave), R = about R rows per CTA
extern int g_reserve_sms;  // SMs left free by the one-wave grids (for a concurrent NCCL kernel when sharded)

struct Plan {
    FusedGeo g;
    long long units;
    int grid;
};

// longest run of output columns, within one kTW-wide tile, that share the same left (x0) or
// right (x1) low-res column: the backward's fast horizontal adjoint handles runs <= kRun
static inline int max_column_run(int w, int W, float sx) {
    int best = 0, run0 = 0, run1 = 0, p0 = -1, p1 = -1;
    for (int x = 0; x < W; ++x) {
        int x0, x1;
        float l0, l1;
        src_index(sx, x, w, x0, x1, l0, l1);
        if (x % kTW == 0) { p0 = p1 = -1; }
        run0 = (x0 == p0) ? run0 + 1 : 1;
        run1 = (x1 == p1) ? run1 + 1 : 1;
        p0 = x0; p1 = x1;
        if (run0 > best) best = run0;
        if (run1 > best) best = run1;
    }
    return best;
}

// geometry + grid for `ctas_per_sm` co-resident CTAs per SM
static inline int make_plan(int C, int h, int w, int H, int W, int n, int ctas_per_sm, Plan& p, int spare_ctas = 0) {
    if (H < h || W < w || H > 65535) return MSQ_E_GEOMETRY;
    FusedGeo& g = p.g;
    g.C = C; g.h = h; g.w = w; g.H = H; g.W = W;
    g.sy = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    g.sx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
    {   // the column-run scan is O(W): remember the last geometries (launches repeat them)
        struct Key { int w, W, run; };
        static Key cache[8];
        static int used = 0, next = 0;
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        int run = -1;
        for (int i = 0; i < used; ++i) if (cache[i].w == w && cache[i].W == W) run = cache[i].run;
        if (run < 0) {
            run = max_column_run(w, W, g.sx);
            cache[next] = Key{w, W, run};
            next = (next + 1) % 8;
            if (used < 8) ++used;
        }
        g.fastx = run <= kRun ? 1 : 0;
    }
    const int tiles_x = (W + kTW - 1) / kTW;
    p.units = (long long)n * tiles_x * H;
    if (p.units >= (1LL << 31)) return MSQ_E_GEOMETRY;
    const int sms = sm_count() - (g_reserve_sms > 0 && g_reserve_sms < sm_count() ? g_reserve_sms : 0);
    long long grid = (long long)sms * ctas_per_sm - spare_ctas;
    if (grid < 1) grid = 1;
    {   // a grid that is a multiple of the column count (images x column tiles) cuts every column into the same number of
        // CTAs: no CTA straddles two columns (a second tile load and set-up in the middle of its rows).  Taken when it
        // costs at most 1/16 of the CTA slots.
        const long long cols = (long long)n * tiles_x;
        const long long aligned = grid / cols * cols;
        if (aligned > 0 && (grid - aligned) * 16 <= grid) grid = aligned;
    }
    if (g_fused_rows > 0) grid = (p.units + g_fused_rows - 1) / g_fused_rows;
    else if (p.units / grid < 4) grid = p.units / 4;          // tiny problems: at least 4 rows per CTA
    // the per-thread packed buckets hold a 16-bit pixel count and a 48-bit fixed-point sum per class (q <= 1, entropy
    // <= ln 32 per pixel): a CTA must not walk more than 2^14 rows (of one image, a fortiori) whatever the knobs say
    const long long min_grid = (p.units + 16383) / 16384;
    if (grid < min_grid) grid = min_grid;
    if (grid < 1) grid = 1;
    if (grid > p.units) grid = p.units;
    p.grid = (int)grid;
    g.uq = (unsigned)(p.units / grid); g.ur = (unsigned)(p.units % grid);
    const long long zc = (long long)n * C * h * w;
    g.zq = (unsigned)(zc / grid); g.zr = (unsigned)(zc % grid);
    long long rmax = (p.units + grid - 1) / grid;
    if (rmax > H) rmax = H;
    g.R = (int)rmax;
    int nrm = (int)ceilf(g.sy * (float)(rmax - 1)) + 3;
    int ncp = (int)ceilf(g.sx * (float)(kTW - 1)) + 3;
    if (nrm > h) nrm = h;
    if (ncp > w) ncp = w;
    g.nrm = nrm;
    g.ncp = ncp;
    return 0;
}

template <typename K>
static int occupancy(K kernel, size_t smem, int fallback) {
    int occ = 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kTW, smem) != cudaSuccess || occ < 1) occ = fallback;
    return occ;
}

// Launch plans are remembered per (kernel, geometry): the occupancy query, the shared-memory
// opt-in and the column-run scan are done once, not on every launch (they cost more host time
// than the launch itself, which matters for the host-buffer pipeline).
struct LaunchPlan {
    Plan p;
    size_t smem;
};
struct PlanKey {
    const void* kernel;
    int C, h, w, H, W, n, rows, device;      // device: the SM count and the shared-memory opt-in are per device
    bool operator==(const PlanKey& o) const {
        return kernel == o.kernel && C == o.C && h == o.h && w == o.w && H == o.H && W == o.W && n == o.n && rows == o.rows &&
               device == o.device;
    }
};
static inline bool plan_cache(const PlanKey& key, LaunchPlan& lp, bool put) {
    constexpr int kSlots = 32;
    static PlanKey keys[kSlots];
    static LaunchPlan vals[kSlots];
    static int used = 0, next = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (put) {
        keys[next] = key; vals[next] = lp;
        next = (next + 1) % kSlots;
        if (used < kSlots) ++used;
        return true;
    }
    for (int i = 0; i < used; ++i) if (keys[i] == key) { lp = vals[i]; return true; }
    return false;
}

// plan + shared-memory size for `kernel`, compiled for `minb` co-resident CTAs per SM.  `spare_ctas`: CTA slots the one-wave
// grid leaves free.  The backward kernels leave one (plus their own extra CTAs): in the two-call path the finalisation kernel's CTA is still resident when they are launched
// (programmatic dependent launch), a full-wave backward then has ONE CTA that cannot be placed until the finalisation exits,
// and on B200 that CTA is often not placed then but only when the first backward CTA exits, 10 us later -- 36.4 instead of
// 31.2 us for such a step (profiles/r02_trace_step.txt, block 2).
template <typename K, typename SmemFn>
static int plan_launch(K kernel, int C, int h, int w, int H, int W, int n, int minb, SmemFn smem_of, LaunchPlan& lp,
                       int spare_ctas = 0) {
    const PlanKey key{(const void*)kernel, C, h, w, H, W, n, g_fused_rows + 100000 * g_reserve_sms + 10000000 * spare_ctas, current_device()};
    if (plan_cache(key, lp, false)) return 0;
    int rc = make_plan(C, h, w, H, W, n, minb, lp.p, spare_ctas);
    if (rc) return rc;
    const int occ = occupancy(kernel, smem_of(lp.p.g), 1);
    if (occ != minb) { rc = make_plan(C, h, w, H, W, n, occ, lp.p, spare_ctas); if (rc) return rc; }
    lp.smem = smem_of(lp.p.g);
    if (lp.smem > 200 * 1024) return MSQ_E_SMEM;
    if (lp.smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.smem);
    plan_cache(key, lp, true);
    return 0;
}

static inline size_t row_tab_bytes(const FusedGeo& g) { return g.R <= kRowTabMax ? (size_t)g.R * 16 : 0; }
static inline size_t tile_bytes(const FusedGeo& g, int ct) { return (size_t)cpd(ct) * g.nrm * g.ncp * sizeof(float); }
static inline size_t fwd_smem(const FusedGeo& g, bool iw, int ct) {
    (void)iw;
    return row_tab_bytes(g) + tile_bytes(g, ct);
}
static inline size_t bwd_smem(const FusedGeo& g, int ct) {
    return row_tab_bytes(g) + tile_bytes(g, ct) +
           ((size_t)(kTW + kRun) * (cpd(ct) + 2) + (size_t)g.ncp * 2 * kRun + 2 * kTW) * sizeof(float) +
           (2 * kTW + 4 * (size_t)g.ncp) * sizeof(int);
}


}  // namespace msq
