// Host-buffer pipeline: the fused loss step (H2D of the head logits, forward, finalise,
// backward, D2H of the loss / histogram / dL/dlogits) for callers whose tensors live in
// HOST memory, software-pipelined over THREE streams -- host->device copies, kernels, device->host
// copies -- with `depth` submissions (slots of device buffers) in flight: the kernels of all steps
// run back to back on one stream exactly as in the device-resident loop (their programmatic
// dependent launches intact, no two one-wave grids competing for the SMs), while the copy engines
// work on the neighbouring steps.  This is the non-PyTorch way into the hot
// path (tools/solve_gta5.py:366-371,199,217 do the same sequence with torch ops: x.to(device),
// model head, loss, backward, .item()).
#include <new>
#include "common.cuh"

using namespace msq;

struct msq_pipe {
    int mode, n, C, h, w, H, W, depth;
    double ratio;
    size_t lo_bytes;
    msq_state_layout lay;
    cudaStream_t s_h2d, s_comp, s_d2h;
    struct Slot {
        cudaEvent_t staged, computed, done;      // input on the device / kernels finished / outputs on the host
        float* d_logits;
        float* d_grad;
        unsigned char* d_accum;
        unsigned char* d_out;
        unsigned char* d_aux;
        bool busy;
    } * slots;
    unsigned long long submitted;
    int n_norm;                 // normaliser when the batch is sharded by image over ranks (0: this pipeline's own n)
    msq_comm* comm;             // statistics exchange of the sharded step (NULL: none)
};

static void pipe_free(msq_pipe* p) {
    if (!p) return;
    if (p->s_h2d) cudaStreamSynchronize(p->s_h2d);
    if (p->s_comp) cudaStreamSynchronize(p->s_comp);
    if (p->s_d2h) cudaStreamSynchronize(p->s_d2h);
    if (p->slots) {
        for (int i = 0; i < p->depth; ++i) {
            msq_pipe::Slot& s = p->slots[i];
            if (s.d_logits) cudaFree(s.d_logits);
            if (s.d_grad) cudaFree(s.d_grad);
            if (s.d_accum) cudaFree(s.d_accum);
            if (s.d_out) cudaFree(s.d_out);
            if (s.d_aux) cudaFree(s.d_aux);
            if (s.staged) cudaEventDestroy(s.staged);
            if (s.computed) cudaEventDestroy(s.computed);
            if (s.done) cudaEventDestroy(s.done);
        }
        delete[] p->slots;
    }
    if (p->s_h2d) cudaStreamDestroy(p->s_h2d);
    if (p->s_comp) cudaStreamDestroy(p->s_comp);
    if (p->s_d2h) cudaStreamDestroy(p->s_d2h);
    delete p;
}

extern "C" int msq_pipe_create(int mode, int n, int num_class, int h, int w, int out_h, int out_w, double ratio,
                               int depth, msq_pipe** out) {
    if (!out || n < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES || h < 1 || w < 1 || out_h < h || out_w < w ||
        depth < 1 || depth > 16)
        return MSQ_E_BADARG;
    if (mode != MSQ_MODE_IW && mode != MSQ_MODE_MAXSQUARE) return MSQ_E_BADARG;
    msq_pipe* p = new (std::nothrow) msq_pipe();
    if (!p) return (int)cudaErrorMemoryAllocation;
    p->mode = mode; p->n = n; p->C = num_class; p->h = h; p->w = w; p->H = out_h; p->W = out_w;
    p->depth = depth; p->ratio = ratio; p->submitted = 0;
    p->n_norm = 0; p->comm = nullptr;
    p->lo_bytes = (size_t)n * num_class * h * w * sizeof(float);
    p->lay = make_layout(n, num_class);
    p->slots = new (std::nothrow) msq_pipe::Slot[depth]();
    if (!p->slots) { delete p; return (int)cudaErrorMemoryAllocation; }
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_comp, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking);
    for (int i = 0; i < depth && e == cudaSuccess; ++i) {
        msq_pipe::Slot& s = p->slots[i];
        if ((e = cudaEventCreateWithFlags(&s.staged, cudaEventDisableTiming)) != cudaSuccess) break;
        if ((e = cudaEventCreateWithFlags(&s.computed, cudaEventDisableTiming)) != cudaSuccess) break;
        if ((e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess) break;
        if ((e = cudaMalloc(&s.d_logits, p->lo_bytes)) != cudaSuccess) break;
        if ((e = cudaMalloc(&s.d_grad, p->lo_bytes)) != cudaSuccess) break;
        if ((e = cudaMalloc(&s.d_accum, (size_t)p->lay.accum_bytes)) != cudaSuccess) break;
        if ((e = cudaMalloc(&s.d_out, (size_t)p->lay.out_bytes)) != cudaSuccess) break;
        if ((e = cudaMalloc(&s.d_aux, (size_t)(16 * (long long)n * out_h * out_w))) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(s.d_accum, 0, (size_t)p->lay.accum_bytes, p->s_comp)) != cudaSuccess) break;
    }
    if (e != cudaSuccess) { pipe_free(p); return (int)e; }
    *out = p;
    return 0;
}

extern "C" int msq_pipe_wait(msq_pipe* p, int slot) {
    if (!p || slot < 0 || slot >= p->depth) return MSQ_E_BADARG;
    msq_pipe::Slot& s = p->slots[slot];
    if (!s.busy) return 0;
    const cudaError_t e = cudaEventSynchronize(s.done);
    s.busy = false;
    return (int)e;
}

// Enqueue one step.  Host buffers should be pinned (cudaHostAlloc / torch pin_memory) for the
// copies to be asynchronous; host_grad / host_hist may be NULL (no backward / no histogram).
// Returns in *slot_out the slot to pass to msq_pipe_wait before reading the outputs or
// reusing the input buffer.  If the slot is still busy the call first waits for it.
extern "C" int msq_pipe_submit(msq_pipe* p, const float* host_logits, float grad_scale, float* host_loss,
                               float* host_grad, int32_t* host_hist, int* slot_out) {
    if (!p || !host_logits || !host_loss) return MSQ_E_BADARG;
    const int slot = (int)(p->submitted % (unsigned long long)p->depth);
    msq_pipe::Slot& s = p->slots[slot];
    int rc = msq_pipe_wait(p, slot);
    if (rc) return rc;
    cudaError_t e;
    // stage 1 (copy engine): input -> device.  The slot's buffers are free: the host waited for its last `done`.
    if ((e = cudaMemcpyAsync(s.d_logits, host_logits, p->lo_bytes, cudaMemcpyHostToDevice, p->s_h2d)) != cudaSuccess) return (int)e;
    if ((e = cudaEventRecord(s.staged, p->s_h2d)) != cudaSuccess) return (int)e;
    // stage 2 (SMs): all steps' kernels in submission order on one stream
    if ((e = cudaStreamWaitEvent(p->s_comp, s.staged, 0)) != cudaSuccess) return (int)e;
    if (host_grad) {
        // the one-call step: forward -> backward (+ finalisation in an extra CTA) and, when the images are sharded over ranks,
        // the exchange of [loss | class histogram] (peer-memory mailboxes in another extra CTA, or ncclAllReduce)
        rc = msq_fused_fwd_bwd(p->mode, s.d_logits, p->n, p->C, p->h, p->w, p->H, p->W, p->ratio, p->n_norm, s.d_accum,
                               s.d_out, s.d_aux, nullptr, grad_scale, s.d_grad, p->comm, 0, (msq_stream_t)p->s_comp);
    } else {
        rc = fused_fwd_dispatch(p->mode, s.d_logits, p->n, p->C, p->h, p->w, p->H, p->W, nullptr, p->ratio, p->n_norm,
                                s.d_accum, s.d_out, nullptr, nullptr, p->s_comp);
    }
    if (rc) return rc;
    if ((e = cudaEventRecord(s.computed, p->s_comp)) != cudaSuccess) return (int)e;
    // stage 3 (the other copy engine): outputs -> host
    if ((e = cudaStreamWaitEvent(p->s_d2h, s.computed, 0)) != cudaSuccess) return (int)e;
    if (host_grad && (e = cudaMemcpyAsync(host_grad, s.d_grad, p->lo_bytes, cudaMemcpyDeviceToHost, p->s_d2h)) != cudaSuccess) return (int)e;
    if ((e = cudaMemcpyAsync(host_loss, s.d_out + p->lay.loss_off, sizeof(float), cudaMemcpyDeviceToHost, p->s_d2h)) != cudaSuccess) return (int)e;
    if (host_hist) {
        e = cudaMemcpyAsync(host_hist, s.d_out + p->lay.hist_out_off, (size_t)p->n * p->C * sizeof(int32_t),
                            cudaMemcpyDeviceToHost, p->s_d2h);
        if (e != cudaSuccess) return (int)e;
    }
    if ((e = cudaEventRecord(s.done, p->s_d2h)) != cudaSuccess) return (int)e;
    s.busy = true;
    p->submitted++;
    if (slot_out) *slot_out = slot;
    return 0;
}

// Images sharded over ranks: every step uses the GLOBAL batch size as the loss normaliser and exchanges its statistics
// vector through `comm` (msq_fused_fwd_bwd).  Call between msq_pipe_create and the first submit, on every rank.
extern "C" int msq_pipe_shard(msq_pipe* p, int n_images_norm, msq_comm* comm) {
    if (!p || n_images_norm < 0 || p->submitted) return MSQ_E_BADARG;
    p->n_norm = n_images_norm;
    p->comm = comm;
    return 0;
}

extern "C" int msq_pipe_drain(msq_pipe* p) {
    if (!p) return MSQ_E_BADARG;
    for (int i = 0; i < p->depth; ++i) {
        const int rc = msq_pipe_wait(p, i);
        if (rc) return rc;
    }
    if (p->comm) return msq_comm_join(p->comm, 0, (msq_stream_t)p->s_comp);       // completes the two steps still in flight
    return 0;
}

extern "C" void msq_pipe_destroy(msq_pipe* p) { pipe_free(p); }
