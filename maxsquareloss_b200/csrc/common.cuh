// Shared device helpers for libmsq_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/msq_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmsq_b200 targets sm_100a (B200) only"
#endif

namespace msq {

constexpr int kSMs = 148;                       // B200: 2 dies x 74 SMs (fallback only: sm_count() asks the device)
constexpr float kFix = 4294967296.0f;           // 2^32: fixed-point scale of q = sum_c p_c^2
constexpr double kInvFix = 1.0 / 4294967296.0;
constexpr unsigned kFlagNonFinite = 1u;
// The per-image accumulators are replicated kRep times (CTA b adds into replica b % kRep):
// a few hundred CTAs finishing together would otherwise queue on the same 2*N*C L2
// addresses (same-address L2 atomics serialise at ~10 ns each on this chip).
constexpr int kRep = 16;

// Carved view of the caller's state buffer (see msq_state_layout in the header).
struct State {
    unsigned int* hist;
    unsigned long long* sumsq;
    unsigned long long* kept;
    unsigned int* flags;
    unsigned int* ticket;
    float* loss;
    float* weights;
    int* hist_out;
    double* sum_out;
    unsigned long long* kept_out;
    double* stats;
    // multi-level guidance (guidance.cu)
    unsigned long long* ce;        // accum: [kRep] sum of -log p2[label_2] in 2^-32 fixed point
    unsigned long long* nvalid;    // accum: [kRep] number of pixels with label_2 != -1
    float* loss2;                  // out
    unsigned long long* nvalid_out;
    double* ce_out;                // out: sum of -log p2[label_2] over the valid pixels (for sharded means)
    unsigned long long* ce_fix_out;   // out: the same sum as the 2^-32 fixed-point integer it was accumulated in (exact to all-reduce)
};

// SM count of the CURRENT device (a process may drive several GPUs: cached per ordinal)
inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    return dev;
}
inline int sm_count() {
    constexpr int kMaxDev = 64;
    static int cache[kMaxDev] = {0};
    const int dev = current_device();
    if (dev < 0 || dev >= kMaxDev) return kSMs;
    int n = cache[dev];
    if (!n) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kSMs;
        cache[dev] = n;
    }
    return n;
}

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

inline msq_state_layout make_layout(int n, int c) {
    msq_state_layout L;
    int64_t nc = (int64_t)n * c, off = 0;
    L.sumsq_off = off;  off += nc * 8 * kRep;          // 8-byte members first
    L.kept_off = off;   off += 8;
    L.hist_off = off;   off += nc * 4 * kRep;
    L.flags_off = off;  off += 4;
    L.ticket_off = off; off += 4;
    off = align_up(off, 8);
    L.ce_off = off;     off += 8 * kRep;
    L.nvalid_off = off; off += 8 * kRep;
    L.accum_bytes = align_up(off, 16);
    off = 0;
    L.sum_out_off = off;  off += (int64_t)n * 8;
    L.kept_out_off = off; off += 8;
    L.loss_off = off;     off += 4;
    off = align_up(off, 16);
    L.weights_off = off;  off += nc * 4;
    L.hist_out_off = off; off += nc * 4;
    off = align_up(off, 16);
    L.stats_off = off;    off += (int64_t)(1 + c) * 8;
    off = align_up(off, 16);
    L.ce_fix_out_off = off; off += 8;      // [ce_fix_out | nvalid_out]: one contiguous uint64 pair (all-reduced as such)
    L.nvalid_out_off = off; off += 8;
    L.loss2_off = off;    off += 4;
    off = align_up(off, 8);
    L.ce_out_off = off;   off += 8;
    L.out_bytes = align_up(off, 16);
    return L;
}

inline State carve(void* accum, void* out, int n, int c) {
    msq_state_layout L = make_layout(n, c);
    char* a = (char*)accum;
    char* o = (char*)out;
    State s;
    s.hist = (unsigned int*)(a + L.hist_off);
    s.sumsq = (unsigned long long*)(a + L.sumsq_off);
    s.kept = (unsigned long long*)(a + L.kept_off);
    s.flags = (unsigned int*)(a + L.flags_off);
    s.ticket = (unsigned int*)(a + L.ticket_off);
    s.loss = (float*)(o + L.loss_off);
    s.weights = (float*)(o + L.weights_off);
    s.hist_out = (int*)(o + L.hist_out_off);
    s.sum_out = (double*)(o + L.sum_out_off);
    s.kept_out = (unsigned long long*)(o + L.kept_out_off);
    s.stats = (double*)(o + L.stats_off);
    s.ce = (unsigned long long*)(a + L.ce_off);
    s.nvalid = (unsigned long long*)(a + L.nvalid_off);
    s.loss2 = (float*)(o + L.loss2_off);
    s.nvalid_out = (unsigned long long*)(o + L.nvalid_out_off);
    s.ce_out = (double*)(o + L.ce_out_off);
    s.ce_fix_out = (unsigned long long*)(o + L.ce_fix_out_off);
    return s;
}

// ---- streaming 128-bit loads/stores: read-once data must not pollute L1 -------------
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ longlong2 ldg_stream_l2(const int64_t* p) {
    longlong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ long long ldg_stream_l1(const int64_t* p) {
    long long r;
    asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}


// Development aid (-DMSQ_TRACE=1, scripts/trace_step.py): thread 0 of every CTA stamps %globaltimer at a few points of the
// step's kernels into a per-translation-unit device array, read back with msq_debug_trace_*().  Off in the product build.
#ifndef MSQ_TRACE
#define MSQ_TRACE 0
#endif
#if MSQ_TRACE
constexpr int kTraceCtas = 1024, kTracePts = 8;
static __device__ unsigned long long g_trace[4 * kTraceCtas * kTracePts];      // [kernel * 2 + step parity][CTA][point]
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smid() {
    unsigned v;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
    return v;
}
#define MSQ_TRACE_PT(kid, i)                                                                           \
    do {                                                                                               \
        if (threadIdx.x == 0 && blockIdx.x < kTraceCtas) {                                             \
            g_trace[((kid) * kTraceCtas + blockIdx.x) * kTracePts + (i)] = gtime();                    \
            if ((i) == 0) g_trace[((kid) * kTraceCtas + blockIdx.x) * kTracePts + 7] = smid();         \
        }                                                                                              \
    } while (0)
#else
#define MSQ_TRACE_PT(kid, i) do {} while (0)
#endif
// stamps of finalize_body: only the finalisation KERNEL's translation unit (api.cu) records them -- inlined into the backward
// (fin_cta) they would land in a forward CTA's slot of that unit's array
#if MSQ_TRACE && defined(MSQ_TRACE_FINALIZE_BODY)
#define MSQ_TRACE_FIN(i) MSQ_TRACE_PT(0, i)
#else
#define MSQ_TRACE_FIN(i) do {} while (0)
#endif

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-
// serialisation attribute may start while its predecessor is still running; it must call
// pdl_wait() before touching anything the predecessor (transitively: any earlier kernel)
// produced.  pdl_trigger() lets the successor's CTAs be scheduled as soon as SM resources
// free up.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// q (or a short run-sum of q's) -> 2^-32 fixed point.  Exact for finite q >= 2^-8:
// scaling by 2^32 is exact and the float then holds an integer < 2^64.
__device__ __forceinline__ unsigned long long to_fix(float qsum) {
    return __float2ull_rn(qsum * kFix);
}

// Image-wise class weight, utils/loss.py:95, with the reference's fp32 roundings:
//   1 / max( fl(hist^r) * fl(total^(1-r)), 1 )     (r and 1-r cast to fp32, as
//   torch.pow(float32 tensor, python scalar) does).  powf is CUDA's 2-ulp fp32 pow
//   (torch's CPU pow is Sleef's 1-ulp one): the weights agree to a few ulp, the loss to ~3e-7.
__device__ __forceinline__ float iw_weight(float hist, float total, float r32, float omr32) {
    const float a = powf(hist, r32);
    const float b = powf(total, omr32);
    const float m = fmaxf(__fmul_rn(a, b), 1.0f);
    return __fdiv_rn(1.0f, m);
}

// Finalisation (the body of api.cu's finalize_kernel, also run by the last CTA of the fused step kernel):
// one warp per image, one lane per class (C <= 32)
// (any block size that is a multiple of 32, up to 256 threads; called by every thread of ONE CTA)
// `defer_clean`: leave the accumulators as they are (finalize_clean() zeroes them later): the one-call step runs this body
// in an extra CTA of its backward kernel while the other CTAs are still reading the class histograms.
__device__ __forceinline__ void finalize_body(const State& st, int mode, int n, int C, float r32, float omr32, int n_norm,
                                              unsigned long long kept_dense, int multi, int loss_kind, bool defer_clean = false) {
    __shared__ double s_red[8];
    __shared__ unsigned long long s_cls[MSQ_MAX_CLASSES];
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
    // thread 0's scalars are fetched first so that their latency overlaps everything else
    unsigned long long kept_ld = 0ull;
    unsigned flags_ld = 0u;
    if (tid == 0) { kept_ld = __ldcg(st.kept); flags_ld = __ldcg(st.flags); }
    // multi-level guidance: cross-entropy sum and valid-pixel count, one replica per lane of warp 0
    unsigned long long ce_ld = 0ull, nv_ld = 0ull;
    if (multi && tid < kRep) {
        ce_ld = st.ce[tid]; nv_ld = st.nvalid[tid];
        if (!defer_clean) { st.ce[tid] = 0ull; st.nvalid[tid] = 0ull; }
    }
    if (tid < MSQ_MAX_CLASSES) s_cls[tid] = 0ull;
    __syncthreads();
    const int nc = n * C;
    double part = 0.0;
    unsigned long long cls_tot = 0ull;
    for (int img = wid; img < n; img += nw) {
        const int idx = img * C + lane;
        unsigned hcnt = 0u;
        unsigned long long sq = 0ull;
        double sd = 0.0;
        if (lane < C) {
            // all 2 x kRep loads are issued before the first use: ONE L2 round trip (the compiler used to start summing after
            // the first 20 and paid a second one: 1.4 us of this kernel's 2.9 us, profiles/r02_trace_step.txt)
            unsigned hv[kRep];
            unsigned long long sv[kRep];
#pragma unroll
            for (int r = 0; r < kRep; ++r) hv[r] = __ldcg(&st.hist[r * nc + idx]);          // L2: other CTAs' atomics
#pragma unroll
            for (int r = 0; r < kRep; ++r) sv[r] = __ldcg(&st.sumsq[r * nc + idx]);
#pragma unroll
            for (int r = 0; r < kRep; ++r) hcnt += hv[r];
            if (loss_kind == 2) {                     // strict MinEnt (softce.cu): fp64 sums, replica order
#pragma unroll
                for (int r = 0; r < kRep; ++r) sd += __longlong_as_double((long long)sv[r]);
            } else {
#pragma unroll
                for (int r = 0; r < kRep; ++r) sq += sv[r];
            }
            if (!defer_clean) {
#pragma unroll
                for (int r = 0; r < kRep; ++r) { st.hist[r * nc + idx] = 0u; st.sumsq[r * nc + idx] = 0ull; }   // self-clean
            }
        }
        const unsigned total = __reduce_add_sync(0xffffffffu, hcnt);
        MSQ_TRACE_FIN(3);
        const double S = (loss_kind == 2) ? sd : (double)sq * kInvFix;
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
    MSQ_TRACE_FIN(4);
    if (lane == 0) s_red[wid] = part;
    __syncthreads();
    MSQ_TRACE_FIN(5);
    if (tid < C) st.stats[1 + tid] = (double)s_cls[tid];
    if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < nw; ++i) tot += s_red[i];
        const unsigned long long kept_local = (kept_dense != 0ull) ? kept_dense : kept_ld;
        const double scale = (double)n_norm / (double)n;
        double loss;
        if (loss_kind == 0) {          // maximum squares: -sum(w p^2)/(N C)   |   -mean(p^2)/2
            if (mode == MSQ_MODE_IW) loss = -tot / ((double)n_norm * (double)C);
            else loss = -tot / (2.0 * (double)kept_local * scale);
        } else {                       // entropy (utils/loss.py:32,65): sum(w H)/(N C)   |   mean(-p log p)
            if (mode == MSQ_MODE_IW) loss = tot / ((double)n_norm * (double)C);
            else loss = tot / ((double)kept_local * scale);
        }
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
            *st.ce_fix_out = ce_ld;
        }
        if (!defer_clean) {
            *st.kept = 0ull;                   // self-clean
            *st.flags = 0u;
        }
        MSQ_TRACE_FIN(6);
    }
}

// the self-clean finalize_body(defer_clean = true) left out; same thread layout
__device__ __forceinline__ void finalize_clean(const State& st, int n, int C, int multi) {
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
    const int nc = n * C;
    if (multi && tid < kRep) { st.ce[tid] = 0ull; st.nvalid[tid] = 0ull; }
    for (int img = wid; img < n; img += nw) {
        const int idx = img * C + lane;
        if (lane < C) {
#pragma unroll
            for (int r = 0; r < kRep; ++r) { st.hist[r * nc + idx] = 0u; st.sumsq[r * nc + idx] = 0ull; }
        }
    }
    if (tid == 0) { *st.kept = 0ull; *st.flags = 0u; }
}


// ---- statistics exchange over NVLink peer memory (comm.cu) -----------------------------------------------
// When the images are sharded over the GPUs of one box, the only data that crosses GPUs is the packed fp64
// statistics vector [loss | class histogram] of each step (<= 264 bytes).  Instead of a collective library call
// per step, the step's own kernels carry the exchange in an extra CTA (of the backward in the one-call step, of the
// finalisation kernel otherwise) that runs beside the finalisation proper: one warp PUSHES the vector this rank produced in the PREVIOUS step into every rank's mailbox
// with 16-byte stores over NVLink (cudaIpc-mapped peer memory), then sums the vectors all ranks pushed one step
// earlier still (they arrived a whole step ago) in rank order and writes that step's all-reduced result.  No
// extra launch, no stream operation between the step's kernels (which would break their programmatic dependent
// launches), no host cost, nothing added to the step's critical path and not one
// register to the hot kernels; the all-reduced vector of step i exists once the finalisation of step i+2 (or
// msq_comm_join) has run.
// Cells are NCCL-LL style {lo32, flag, hi32, flag}: data and flag travel in the same 8-byte store, so the
// reader needs no fence: it polls until both flags carry the sequence number it expects.
constexpr int kMaxPeers = 8;       // GPUs of one NVSwitch box
constexpr int kBoxSlots = 8;       // ring of sequence numbers: when vector s is pushed, vectors s-3 .. s-1 may still be read (>= 4 slots)
constexpr int kBoxCount = 40;      // doubles per vector (1 + 32 classes, rounded up)

// Static part, resident in device memory (written once by msq_comm_box_open).
struct PeerBoxStatic {
    int world, rank;
    unsigned* err;                 // error words in MAPPED PINNED HOST memory (the host reads them without synchronising):
                                   //   [0] bit 0 = some peer's vector did not arrive in time, [1] = how many vectors were lost
    unsigned long long timeout_ns; // how long one reduction waits for a peer (msq_comm_box_timeout; default 600 s, the order of
                                   // NCCL's watchdog).  Re-armed for every vector: a late peer costs NaN statistics for the
                                   // vectors it missed, not for the rest of the run
    uint4* peer[kMaxPeers];        // rank p's mailbox: [kBoxSlots][kMaxPeers][kBoxCount] cells (peer[rank] is local)
};
// Per-step part, passed to the kernel by value (st == NULL: no exchange).  Every pointer is into memory the COMMUNICATOR
// owns (msq_comm: box_vec / box_red rings): nothing here outlives or aliases a caller's buffer.
struct PeerBox {
    const PeerBoxStatic* st;
    const double* cur;             // vector to push: this rank's statistics of sequence number `seq` (NULL: nothing to push)
    double* prev_out;              // where the all-reduced vector of sequence number `prev_seq` goes (NULL: nothing to reduce)
    double* keep;                  // where the finalisation proper stores this step's own vector for a later push (NULL: no)
    unsigned seq, prev_seq;        // 1, 2, ...; 0 is never used (it is the flag value of an empty cell)
    short count, prev_count, keep_count;       // doubles per vector
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void ll_store(uint4* cell, double v, unsigned flag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};"
                 :: "l"(cell), "r"((unsigned)b), "r"(flag), "r"((unsigned)(b >> 32)), "r"(flag) : "memory");
}
__device__ __forceinline__ bool ll_load(const uint4* cell, unsigned flag, double& v) {
    unsigned a, fa, b, fb;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(cell) : "memory");
    if (fa != flag || fb != flag) return false;
    v = __longlong_as_double((long long)(((unsigned long long)b << 32) | a));
    return true;
}
__device__ __forceinline__ uint4* box_cell(uint4* base, unsigned seq, int src_rank, int k) {
    return base + ((size_t)(seq % kBoxSlots) * kMaxPeers + src_rank) * kBoxCount + k;
}
// sum over the ranks, in rank order (identical bits on every rank), of the vectors pushed for `seq`.  One warp.
__device__ __forceinline__ void box_reduce(const PeerBoxStatic* st, unsigned seq, int count, double* out, int lane) {
    const int world = st->world, rank = st->rank;
    uint4* mine = st->peer[rank];
    const unsigned long long limit = st->timeout_ns;
    unsigned long long t0 = 0ull;
    bool lost = false;
    for (int k = lane; k < count; k += 32) {
        double sum = 0.0;
        for (int p = 0; p < world; ++p) {
            const uint4* cell = box_cell(mine, seq, p, k);
            double v = 0.0;
            unsigned spins = 0;
            while (!lost && !ll_load(cell, seq, v)) {
                if ((++spins & 255u) == 0u) {                     // look at the clock every 256 polls
                    const unsigned long long now = global_ns();
                    if (t0 == 0ull) t0 = now;
                    else if (now - t0 > limit) lost = true;       // this vector is given up; the next one waits afresh
                }
                __nanosleep(64);
            }
            if (lost) v = __longlong_as_double(0x7ff8000000000000LL);
            sum += v;
        }
        out[k] = sum;
    }
    if (__any_sync(0xffffffffu, lost) && lane == 0) {             // one warp per GPU ever writes these words
        volatile unsigned* e = st->err;
        e[0] = e[0] | 1u;
        e[1] = e[1] + 1u;
    }
}
// One warp: push vector `seq` to every rank (itself included), then reduce vector `prev_seq`.
__device__ __forceinline__ void box_exchange(const PeerBox& b, int lane) {
    const PeerBoxStatic* st = b.st;
    if (b.cur) {
        const int world = st->world, rank = st->rank;
        for (int k = lane; k < b.count; k += 32) {
            const double v = b.cur[k];
            for (int p = 0; p < world; ++p) ll_store(box_cell(st->peer[p], b.seq, rank, k), v, b.seq);
        }
    }
    if (b.prev_out) box_reduce(st, b.prev_seq, b.prev_count, b.prev_out, lane);
}

// The finalisation's arguments as ONE kernel argument of the fused backward (one-call step: `extra` CTAs past the work grid
// run it, fused_common.cuh); extra == 0: the backward is an ordinary one (weights from `out`, finalisation kernel before it)
struct FinArgs {
    State st;
    PeerBox box;
    unsigned long long kept_dense;
    int mode, n, C, n_norm, loss_kind, extra;
    float r32, omr32;
};

// Finalisation kernel (api.cu), launched on the same stream right after a forward kernel:
// sums the replicas, turns the integer accumulators into the reference's scalar, the
// per-image weights and the final histogram, and re-zeroes the accumulators so the
// buffer is ready for the next call without a memset.
//   IW        (utils/loss.py:95,100):  loss = -(1/(Nn*C)) sum_n sum_k w_nk * S_nk
//   MaxSquare (utils/loss.py:118):     loss = -(sum q) / (2 * kept)
// Nn = n_norm (global batch when sharded).  All sums in fp64, fixed order.
int launch_finalize(const State& st, int mode, int n, int C, float r32, float omr32, int n_norm,
                    unsigned long long kept_dense, cudaStream_t stream, int multi = 0, int loss_kind = 0,
                    const PeerBox* box = nullptr);

// fused_loss.cu entry points shared with the host pipeline (host_pipe.cu)
int fused_fwd_dispatch(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                       const int64_t* label, double ratio, int n_images_norm, void* accum, void* out, void* aux,
                       float* zero_grad, cudaStream_t s, int loss_kind = 0, const PeerBox* box = nullptr, int late_finalize = 0);
int fused_bwd_dispatch(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                       int n_images_norm, const void* out, const float* grad_out, float grad_out_value,
                       float* grad_logits, const void* aux, int grad_is_zeroed, cudaStream_t s, int loss_kind = 0,
                       void* accum_fin = nullptr, double ratio = 0.0, const PeerBox* box = nullptr);
extern int g_late_finalize;    // fused_loss.cu, tuning knob "late_finalize" (default 1)

}  // namespace msq

namespace msq {
// every kernel this library launches is counted (msq_launch_count: bench.py reports the count of its timed region)
extern unsigned long long g_launches;
inline void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
// launch with the programmatic-stream-serialisation attribute (see pdl_wait / pdl_trigger)
extern int g_pdl_mask;      // api.cu, tuning knob "pdl_mask": bit 0 forward kernels, bit 1 finalisation, bit 2 backward kernels, bit 3 the rest
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_as(int kind_bit, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (g_pdl_mask & kind_bit) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    if (e == cudaSuccess) count_launch();
    return e;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    return launch_pdl_as(8, kernel, grid, block, smem, stream, args...);
}
}  // namespace msq

#define MSQ_CHECK_LAUNCH()                       \
    do {                                         \
        cudaError_t e__ = cudaGetLastError();    \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)
