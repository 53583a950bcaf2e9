// Host binding of the loss entry points for PyTorch callers: torch::autograd::Function nodes written in C++ that call
// the C ABI of libmsq_b200.so (include/msq_b200.h) -- the same symbols the ctypes binding (_lib.py) calls.
//
// Why native: the reference's call sites (tools/solve_gta5.py:199,217)
//     loss = self.target_loss(pred, pred_P);  (lambda_target * loss).backward()
// cost ~140 us of host time per step through a Python autograd.Function (Function.apply, ctx, the hop of the Python
// backward into the engine's device thread), four times the 34 us of GPU work they enqueue.  A C++ node leaves only
// PyTorch's own fixed costs (the caller's `lambda * loss` and the engine start-up).  Nothing here computes anything:
// PyTorch supplies the device memory, the current stream and the autograd graph; the kernels are the library's.
//
// Registered as torch.ops.msq_b200.{fused_loss, prob_loss}; loaded by maxsquareloss_b200/_torch_ops.py.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/autograd.h>
#include <torch/library.h>

#include <mutex>
#include <unordered_map>

#include "../../include/msq_b200.h"

namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

inline void check_rc(int rc) { TORCH_CHECK(rc == 0, "libmsq_b200: ", msq_error_string(rc), " (code ", rc, ")"); }

const msq_state_layout& layout_of(int n, int c) {
    static std::mutex mu;
    static std::unordered_map<int64_t, msq_state_layout> cache;
    std::lock_guard<std::mutex> g(mu);
    const int64_t key = ((int64_t)n << 8) | c;
    auto it = cache.find(key);
    if (it == cache.end()) {
        msq_state_layout lay;
        check_rc(msq_state_layout_get(n, c, &lay));
        it = cache.emplace(key, lay).first;
    }
    return it->second;
}

// zero-initialised, self-cleaning accumulators: one buffer per (device, stream), reused forever (msq_b200.h, "Loss buffers")
Tensor accum_of(const c10::Device& dev, cudaStream_t stream, int64_t nbytes) {
    static std::mutex mu;
    static std::unordered_map<uint64_t, Tensor> cache;
    std::lock_guard<std::mutex> g(mu);
    const uint64_t key = ((uint64_t)(uintptr_t)stream) ^ ((uint64_t)dev.index() << 56);
    auto it = cache.find(key);
    if (it == cache.end() || it->second.numel() < nbytes) {
        Tensor t = at::zeros({std::max<int64_t>(nbytes, 4096)}, at::TensorOptions().dtype(at::kByte).device(dev));
        if (it == cache.end()) it = cache.emplace(key, t).first;
        else it->second = t;
    }
    return it->second;
}

void require_cuda_f32_4d(const Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor: maxsquareloss_b200 has no CPU fallback");
    TORCH_CHECK(t.scalar_type() == at::kFloat, name, " must be float32, got ", t.scalar_type());
    TORCH_CHECK(t.dim() == 4, name, " must be (N,C,H,W), got ", t.dim(), " dimensions");
}

const void* label_ptr(const c10::optional<Tensor>& label, const Tensor& like, int64_t n, int64_t H, int64_t W) {
    if (!label.has_value() || !label->defined()) return nullptr;
    const Tensor& l = *label;
    TORCH_CHECK(l.is_cuda() && l.device() == like.device() && l.scalar_type() == at::kLong && l.is_contiguous(),
                "label must be a contiguous int64 CUDA tensor on the loss's device");
    TORCH_CHECK(l.dim() == 3 && l.size(0) == n && l.size(1) == H && l.size(2) == W, "label must have shape (N,H,W)");
    return l.data_ptr();
}

Tensor grad_scalar(const Tensor& go, const Tensor& like) {
    if (go.device() == like.device() && go.scalar_type() == at::kFloat && go.is_contiguous()) return go;
    return go.to(like.device(), at::kFloat).contiguous();
}

// ---- fused: low-resolution head logits (K1/K2; kind 1 = the MinEnt variant) --------------------------------------
struct FusedLossFn : public torch::autograd::Function<FusedLossFn> {
    static variable_list forward(AutogradContext* ctx, const Tensor& logits, const c10::optional<Tensor>& label, int64_t H,
                                 int64_t W, int64_t mode, double ratio, int64_t n_norm, int64_t kind, bool need_grad,
                                 bool use_cache) {
        require_cuda_f32_4d(logits, "head logits");
        const Tensor lo = logits.contiguous();
        const int64_t n = lo.size(0), c = lo.size(1), h = lo.size(2), w = lo.size(3);
        TORCH_CHECK(c <= MSQ_MAX_CLASSES, "num_class=", c, " exceeds the kernels' limit of ", MSQ_MAX_CLASSES);
        const c10::cuda::CUDAGuard guard(lo.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream(lo.device().index()).stream();
        const msq_state_layout& lay = layout_of((int)n, (int)c);
        Tensor accum = accum_of(lo.device(), stream, lay.accum_bytes);
        Tensor out = at::empty({lay.out_bytes >> 2}, lo.options());
        Tensor aux, grad;
        const bool cache = need_grad && (use_cache || kind == 1);
        if (cache || kind == 1) aux = at::empty({16 * n * H * W}, lo.options().dtype(at::kByte));      // msq_fused_aux_bytes
        if (cache) grad = at::empty_like(lo);
        const void* lab = label_ptr(label, lo, n, H, W);
        int rc;
        if (kind == 0)
            rc = msq_fused_fwd((int)mode, lo.data_ptr<float>(), (int)n, (int)c, (int)h, (int)w, (int)H, (int)W,
                               (const int64_t*)lab, ratio, (int)n_norm, accum.data_ptr(), out.data_ptr(),
                               aux.defined() ? aux.data_ptr() : nullptr, grad.defined() ? grad.data_ptr<float>() : nullptr,
                               (msq_stream_t)stream);
        else
            rc = msq_entropy_fwd((int)mode, lo.data_ptr<float>(), (int)n, (int)c, (int)h, (int)w, (int)H, (int)W, ratio,
                                 (int)n_norm, accum.data_ptr(), out.data_ptr(), aux.data_ptr(),
                                 grad.defined() ? grad.data_ptr<float>() : nullptr, (msq_stream_t)stream);
        check_rc(rc);
        Tensor loss = out.select(0, lay.loss_off >> 2);
        if (need_grad) {
            ctx->save_for_backward({lo});
            ctx->saved_data["out"] = out;
            if (aux.defined()) ctx->saved_data["aux"] = aux;
            if (grad.defined()) ctx->saved_data["grad"] = grad;
            ctx->saved_data["cfg"] = std::vector<int64_t>{mode, H, W, n_norm, kind};
        }
        ctx->mark_non_differentiable({out});
        return {loss, out};
    }

    static variable_list backward(AutogradContext* ctx, variable_list grad_outputs) {
        variable_list res(10);
        if (!ctx->needs_input_grad(0) || !grad_outputs[0].defined()) return res;
        const Tensor lo = ctx->get_saved_variables()[0];
        const auto cfg = ctx->saved_data["cfg"].toIntVector();
        const int64_t mode = cfg[0], H = cfg[1], W = cfg[2], n_norm = cfg[3], kind = cfg[4];
        const Tensor out = ctx->saved_data["out"].toTensor();
        Tensor aux, grad;
        if (ctx->saved_data.count("aux")) aux = ctx->saved_data["aux"].toTensor();
        int zeroed = 0;
        if (ctx->saved_data.count("grad")) {          // the buffer the forward zero-filled is good for ONE backward
            grad = ctx->saved_data["grad"].toTensor();
            ctx->saved_data.erase("grad");
            zeroed = 1;
        } else {
            grad = at::empty_like(lo);
        }
        const Tensor go = grad_scalar(grad_outputs[0], lo);
        const c10::cuda::CUDAGuard guard(lo.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream(lo.device().index()).stream();
        const int n = (int)lo.size(0), c = (int)lo.size(1), h = (int)lo.size(2), w = (int)lo.size(3);
        int rc;
        if (kind == 0)
            rc = msq_fused_bwd((int)mode, lo.data_ptr<float>(), n, c, h, w, (int)H, (int)W, (int)n_norm, out.data_ptr(),
                               aux.defined() ? aux.data_ptr() : nullptr, go.data_ptr<float>(), grad.data_ptr<float>(), zeroed,
                               (msq_stream_t)stream);
        else
            rc = msq_entropy_bwd((int)mode, lo.data_ptr<float>(), n, c, h, w, (int)H, (int)W, (int)n_norm, out.data_ptr(),
                                 aux.data_ptr(), go.data_ptr<float>(), grad.data_ptr<float>(), zeroed, (msq_stream_t)stream);
        check_rc(rc);
        res[0] = grad;
        return res;
    }
};

// ---- strict drop-in: full-resolution probabilities (K3/K4) --------------------------------------------------------
struct ProbLossFn : public torch::autograd::Function<ProbLossFn> {
    static variable_list forward(AutogradContext* ctx, const Tensor& prob, const c10::optional<Tensor>& label, int64_t mode,
                                 double ratio, int64_t ignore_index, int64_t n_norm, bool need_grad) {
        require_cuda_f32_4d(prob, "prob");
        const Tensor p = prob.contiguous();
        const int64_t n = p.size(0), c = p.size(1), H = p.size(2), W = p.size(3);
        TORCH_CHECK(c <= MSQ_MAX_CLASSES, "num_class=", c, " exceeds the kernels' limit of ", MSQ_MAX_CLASSES);
        const c10::cuda::CUDAGuard guard(p.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream(p.device().index()).stream();
        const msq_state_layout& lay = layout_of((int)n, (int)c);
        Tensor accum = accum_of(p.device(), stream, lay.accum_bytes);
        Tensor out = at::empty({lay.out_bytes >> 2}, p.options());
        const void* lab = label_ptr(label, p, n, H, W);
        check_rc(msq_prob_fwd((int)mode, p.data_ptr<float>(), (int)n, (int)c, H * W, (const int64_t*)lab, ratio,
                              (int)ignore_index, (int)n_norm, accum.data_ptr(), out.data_ptr(), (msq_stream_t)stream));
        if (need_grad) {
            ctx->save_for_backward({p});
            ctx->saved_data["out"] = out;
            ctx->saved_data["cfg"] = std::vector<int64_t>{mode, ignore_index, n_norm};
        }
        ctx->mark_non_differentiable({out});
        return {out.select(0, lay.loss_off >> 2), out};
    }

    static variable_list backward(AutogradContext* ctx, variable_list grad_outputs) {
        variable_list res(7);
        if (!ctx->needs_input_grad(0) || !grad_outputs[0].defined()) return res;
        const Tensor p = ctx->get_saved_variables()[0];
        const auto cfg = ctx->saved_data["cfg"].toIntVector();
        const Tensor out = ctx->saved_data["out"].toTensor();
        const Tensor go = grad_scalar(grad_outputs[0], p);
        Tensor grad = at::empty_like(p);
        const c10::cuda::CUDAGuard guard(p.device());
        cudaStream_t stream = at::cuda::getCurrentCUDAStream(p.device().index()).stream();
        check_rc(msq_prob_bwd((int)cfg[0], p.data_ptr<float>(), (int)p.size(0), (int)p.size(1), p.size(2) * p.size(3),
                              (int)cfg[1], (int)cfg[2], out.data_ptr(), go.data_ptr<float>(), grad.data_ptr<float>(),
                              (msq_stream_t)stream));
        res[0] = grad;
        return res;
    }
};

std::tuple<Tensor, Tensor> fused_loss(const Tensor& logits, const c10::optional<Tensor>& label, int64_t H, int64_t W,
                                      int64_t mode, double ratio, int64_t n_norm, int64_t kind, bool use_cache) {
    const bool need = at::GradMode::is_enabled() && logits.requires_grad();
    auto r = FusedLossFn::apply(logits, label, H, W, mode, ratio, n_norm, kind, need, use_cache);
    return {r[0], r[1]};
}

std::tuple<Tensor, Tensor> prob_loss(const Tensor& prob, const c10::optional<Tensor>& label, int64_t mode, double ratio,
                                     int64_t ignore_index, int64_t n_norm) {
    const bool need = at::GradMode::is_enabled() && prob.requires_grad();
    auto r = ProbLossFn::apply(prob, label, mode, ratio, ignore_index, n_norm, need);
    return {r[0], r[1]};
}

}  // namespace

TORCH_LIBRARY(msq_b200, m) {
    m.def("fused_loss(Tensor logits, Tensor? label, int H, int W, int mode, float ratio, int n_norm, int kind, bool use_cache) -> (Tensor, Tensor)",
          &fused_loss);
    m.def("prob_loss(Tensor prob, Tensor? label, int mode, float ratio, int ignore_index, int n_norm) -> (Tensor, Tensor)",
          &prob_loss);
}
