/*
 * msq_b200.h -- C ABI of libmsq_b200.so: the sm_100a (B200) implementation of
 * MaxSquareLoss's per-pixel adaptation-loss and evaluation hot path.
 *
 * The reference (shiyutang/MaxSquareLoss) is pure Python; it has no FFI.  The
 * boundary it offers is class substitution at its factory sites
 * (tools/solve_gta5.py:156-160, tools/solve_crosscity.py:99-102,
 * tools/train_source.py:125).  Each entry point below is what a binding for
 * that class/method would call; the reference interface it replaces is cited.
 * The ctypes binding is maxsquareloss_b200/_lib.py; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C symbols, no exceptions, no ownership: the caller allocates every
 *     device buffer (torch tensors in the Python shim) and passes raw device
 *     pointers, sizes and the CUDA stream to enqueue on (cudaStream_t as void*).
 *   - every call is asynchronous on `stream` and never synchronises the host.
 *   - return value: 0 on success, a cudaError_t (>0) from a launch / memset, or
 *     an MSQ_E_* code (<0) for argument errors.  msq_error_string() names it.
 *   - tensors are dense row-major ("NCHW contiguous"), fp32; labels int64.
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef MSQ_B200_H_
#define MSQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSQ_ABI_VERSION 6
#define MSQ_MAX_CLASSES 32          /* reference uses 13 / 16 / 19 */

#define MSQ_E_BADARG   (-1)         /* null pointer, non-positive size, C > MSQ_MAX_CLASSES */
#define MSQ_E_GEOMETRY (-2)         /* fused path needs H >= h and W >= w (the model only upsamples) */
#define MSQ_E_SMEM     (-3)         /* tile does not fit shared memory */
#define MSQ_E_NCCL     (-5)         /* libnccl.so.2 could not be loaded, or an NCCL call failed */
#define MSQ_E_ALIGN    (-4)         /* pointer not aligned for its element type */
#define MSQ_E_PEER     (-6)         /* a peer's statistics vector did not arrive within the mailbox time-out */
#define MSQ_E_NOTREADY (-7)         /* msq_comm_result: that step's all-reduced vector has not been produced yet */

#define MSQ_MODE_MAXSQUARE 0        /* utils/loss.py:104-119  MaxSquareloss      */
#define MSQ_MODE_IW        1        /* utils/loss.py:69-102   IW_MaxSquareloss   */

typedef void* msq_stream_t;         /* cudaStream_t */

int         msq_abi_version(void);
const char* msq_error_string(int code);
/* number of kernels this library has launched in this process so far (bench.py reports the count of its timed region;
 * memsets and NCCL's own kernels are not counted) */
unsigned long long msq_launch_count(void);

/* ---------------------------------------------------------------------------
 * Loss buffers.  A forward call uses two device buffers (byte offsets from
 * msq_state_layout_get):
 *
 * ACCUMULATORS (`accum`, accum_bytes): must be ZERO when a *_fwd call starts and
 * are zero again when that call's kernels have finished (the finalisation kernel
 * that every *_fwd call enqueues re-zeroes them), so ONE zero-initialised buffer
 * per stream can be reused forever without a memset.  Not shareable between streams.
 *   sumsq  uint64[R][N*C] sum over pixels of q = sum_c p_c^2 in 2^-32 fixed point,
 *                       bucketed by the pixel's argmax class (IW) or in bucket 0
 *                       of each image (MaxSquare).  Integer, hence
 *                       order-independent: the loss is bit-reproducible.
 *                       R = 16 replicas (CTA b adds into replica b % R) keep
 *                       same-address L2 atomics off the critical path.
 *   kept   uint64       number of prob elements != ignore_index (MaxSquare mean)
 *   hist   uint32[R][N*C] per-image class histogram being accumulated
 *   flags  uint32       bit0: a non-finite value was seen (loss becomes NaN)
 *   ticket uint32       msq_fused_fwd_bwd: work CTAs of the backward that are done with the accumulators (zero between steps)
 *   ce     uint64[R]    multi-level guidance: sum of -log p2[label_2], 2^-32 fixed point
 *   nvalid uint64[R]    multi-level guidance: pixels with label_2 != -1
 *
 * OUTPUTS (`out`, out_bytes): written by the forward, read by the backward.
 *   sum_out   float64[N]   per-image sum of q (diagnostics / multi-GPU reduction)
 *   kept_out  uint64       elements kept by the MaxSquare mask (local)
 *   loss      float32      the scalar the reference's forward() returns
 *   weights   float32[N*C] image-wise class weights (utils/loss.py:95); 1 for MaxSquare
 *   hist_out  int32[N*C]   final per-image histogram (what utils/loss.py:92-94 computes)
 *   stats     float64[1+C] [loss, sum over images of hist_out]: the rank-local
 *                          partials, packed so that ONE NCCL all-reduce(sum) of
 *                          this vector yields the global loss and class histogram
 *                          when images are sharded over ranks (counts < 2^53 are
 *                          exact in fp64)
 *   nvalid_out uint64      msq_multi_fwd: number of pixels with label_2 != -1
 *   loss2     float32      msq_multi_fwd: CrossEntropyLoss(ignore_index=-1)(head 2, label_2)
 *   ce_out    float64      msq_multi_fwd: sum over those pixels of -log softmax(head 2)[label_2];
 *                          loss2 = ce_out / nvalid_out.  When images are sharded over ranks the
 *                          caller all-reduces the pair below in place before msq_guidance_bwd (which
 *                          divides by nvalid_out) and forms loss2 = ce_fix_out * 2^-32 / nvalid_out
 *   ce_fix_out uint64      the same sum as the 2^-32 fixed-point integer it was accumulated in;
 *                          [ce_fix_out, nvalid_out] are ADJACENT: one exact integer all-reduce of 2 words
 *                          (msq_comm_allreduce_u64) gives the cross-entropy mean of the GLOBAL batch
 * ------------------------------------------------------------------------- */
typedef struct msq_state_layout {
    int64_t sumsq_off, kept_off, hist_off, flags_off, ticket_off, ce_off, nvalid_off;   /* into accum */
    int64_t accum_bytes;
    int64_t sum_out_off, kept_out_off, loss_off, weights_off, hist_out_off, stats_off,
            ce_fix_out_off, nvalid_out_off, loss2_off, ce_out_off;                                /* into out */
    int64_t out_bytes;
} msq_state_layout;

int msq_state_layout_get(int n_images, int num_class, msq_state_layout* out);

/* ---------------------------------------------------------------------------
 * Strict drop-in: full-resolution probabilities in, as the reference's
 * forward(pred, prob[, label]) receives them (`pred` is unused by the
 * reference, utils/loss.py:84,117, and is not passed).
 *
 *   mode          MSQ_MODE_MAXSQUARE | MSQ_MODE_IW
 *   prob          float32 [N,C,H*W]
 *   label         int64 [N,H*W] or NULL  (IW only, utils/loss.py:87-88: the map
 *                 whose values are counted; weights are still gathered by argmax)
 *   ratio         IW ratio (utils/loss.py:74), ignored for MaxSquare
 *   ignore_index  utils/loss.py:72,107 (compared against prob as float)
 *   n_images_norm normaliser N of utils/loss.py:100 (pass the GLOBAL batch size
 *                 when images are sharded over ranks; 0 means N)
 *   accum, out    device buffers, see above
 * ------------------------------------------------------------------------- */
int msq_prob_fwd(int mode, const float* prob, int n, int num_class, int64_t hw,
                 const int64_t* label, double ratio, int ignore_index, int n_images_norm,
                 void* accum, void* out, msq_stream_t stream);

/* dL/dprob = grad_out * d(loss)/d(prob), written densely (zeros where masked).
 *   grad_out   device pointer to the 0-dim fp32 upstream gradient (callers
 *              scale the loss by lambda_target before backward,
 *              tools/solve_gta5.py:199,217)
 *   out        the output buffer msq_prob_fwd filled (weights / kept_out are read)   */
int msq_prob_bwd(int mode, const float* prob, int n, int num_class, int64_t hw,
                 int ignore_index, int n_images_norm, const void* out,
                 const float* grad_out, float* grad_prob, msq_stream_t stream);

/* ---------------------------------------------------------------------------
 * Fused: low-resolution head logits in.  Absorbs the model's
 * F.interpolate(..., mode='bilinear', align_corners=True)
 * (graphs/models/deeplab_multi.py:124,128), the trainer's F.softmax(pred, 1)
 * (tools/solve_gta5.py:182-183) and the loss; the backward goes through the
 * softmax and the bilinear adjoint and writes dL/dlogits at h x w.  No
 * full-resolution tensor is read or written.
 *   logits      float32 [N,C,h,w]
 *   label       int64 [N,H,W] or NULL (IW `label=` argument, full resolution)
 *   aux         optional (NULL = off) per-pixel statistics cache of
 *               msq_fused_aux_bytes(N,H,W) = 16*N*H*W bytes (float4 {max, q*s, 1/s^2,
 *               argmax class}): written by the forward, and when handed to the
 *               backward it skips re-deriving them (~55 % fewer instructions).  Not a
 *               probability tensor: 16 B/pixel instead of 4*C.
 *   zero_grad   optional float32 [N,C,h,w] buffer the forward zero-fills on the side
 *               (pass the future grad_logits and grad_is_zeroed=1 to the backward to
 *               save the memset launch)
 *   grad_logits float32 [N,C,h,w], overwritten
 * ------------------------------------------------------------------------- */
int64_t msq_fused_aux_bytes(int n, int out_h, int out_w);

int msq_fused_fwd(int mode, const float* logits, int n, int num_class, int h, int w,
                  int out_h, int out_w, const int64_t* label, double ratio,
                  int n_images_norm, void* accum, void* out, void* aux /* nullable */,
                  float* zero_grad /* nullable */, msq_stream_t stream);

int msq_fused_bwd(int mode, const float* logits, int n, int num_class, int h, int w,
                  int out_h, int out_w, int n_images_norm, const void* out,
                  const void* aux /* nullable */, const float* grad_out, float* grad_logits,
                  int grad_is_zeroed, msq_stream_t stream);

/* MinEnt baselines of the same loss factory (tools/solve_gta5.py:150-155): softCrossEntropy
 * (mode MSQ_MODE_MAXSQUARE = unweighted) and IWsoftCrossEntropy (mode MSQ_MODE_IW) of
 * utils/loss.py:17-67, called as the trainers call them -- target = softmax(inputs), attached to the
 * graph (tools/solve_gta5.py:188-190,199) -- fused from the low-resolution head logits:
 *   loss = mean(-p log p)                                   dL/dz_j = -(1/M) p_j (log p_j + H)
 *   loss = sum_px w[argmax_c inputs] H_px / (N C)            H = -sum_c p_c log p_c
 * (IWsoftCrossEntropy takes the argmax of the LOGITS and a C-bin histc, utils/loss.py:54-60).
 * Buffers and outputs as msq_fused_fwd / msq_fused_bwd; aux is REQUIRED (the backward replays it). */
int msq_entropy_fwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                    double ratio, int n_images_norm, void* accum, void* out, void* aux, float* zero_grad /* nullable */,
                    msq_stream_t stream);
int msq_entropy_bwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                    int n_images_norm, const void* out, const void* aux, const float* grad_out, float* grad_logits,
                    int grad_is_zeroed, msq_stream_t stream);

/* Strict MinEnt: the reference classes on full-resolution tensors and an ARBITRARY target (utils/loss.py:23-35, 46-67):
 *   mode MSQ_MODE_MAXSQUARE  softCrossEntropy:    mean( (-log_softmax(inputs,1) * target)[target != ignore_index] )
 *   mode MSQ_MODE_IW         IWsoftCrossEntropy:  sum( (-log_softmax * target * w[argmax_c inputs])[mask] ) / (N C)
 *   inputs, target  float32 [N,C,H*W]
 * Buffers as msq_prob_fwd (out.hist_out = per-image histogram of argmax(inputs), out.weights, out.loss, out.stats).
 * The backward writes d/d inputs and, when grad_target != NULL, d/d target (the trainers' target = softmax(inputs) is
 * attached to the graph, tools/solve_gta5.py:188-190): both overwritten, zeros where target == ignore_index. */
int msq_softce_fwd(int mode, const float* inputs, const float* target, int n, int num_class, int64_t hw,
                   double ratio, int ignore_index, int n_images_norm, void* accum, void* out, msq_stream_t stream);
int msq_softce_bwd(int mode, const float* inputs, const float* target, int n, int num_class, int64_t hw,
                   int ignore_index, int n_images_norm, const void* out, const float* grad_out,
                   float* grad_inputs, float* grad_target /* nullable */, msq_stream_t stream);

/* ---------------------------------------------------------------------------
 * Multi-level self-produced guidance ("MaxSquare+IW+Multi", BASELINE config 3): the inline
 * trainer code tools/solve_gta5.py:183,192,206-215 == tools/solve_crosscity.py:235-243,
 * fused with the head-1 adaptation loss, from the LOW-resolution outputs of both heads.
 *   loss   (out.loss)  = msq_fused_fwd's loss on head 1 (mode, ratio)
 *   label_2            = (max P1 > threshold || max P2 > threshold) ? argmax_c (P1+P2)/2 : -1
 *   loss2  (out.loss2) = nn.CrossEntropyLoss(ignore_index=-1)(up(head 2), label_2)
 *                        mean over the out.nvalid_out pixels with label_2 != -1 (NaN if none)
 *   aux1 / aux2        float4-per-pixel caches (msq_fused_aux_bytes each, nullable) for
 *                      msq_fused_bwd (head 1) and msq_guidance_bwd (head 2, required there)
 *   zero_grad1/2       optional dL/dlogits buffers to zero-fill on the side
 *   label2_out         optional int64 [N,H,W]: the pseudo-label map (bit-exact vs the reference)
 * The caller combines  lambda_target*loss + lambda_seg*lambda_target*loss2  (solve_gta5.py:199,213).
 * ------------------------------------------------------------------------- */
int msq_multi_fwd(int mode, const float* logits1, const float* logits2, int n, int num_class,
                  int h, int w, int out_h, int out_w, double ratio, double threshold,
                  int n_images_norm, void* accum, void* out, void* aux1, void* aux2,
                  float* zero_grad1, float* zero_grad2, int64_t* label2_out, msq_stream_t stream);

/* d(loss2)/d(logits2) = grad_out/nvalid * (softmax - onehot(label_2)) through the bilinear adjoint. */
int msq_guidance_bwd(const float* logits2, int n, int num_class, int h, int w, int out_h, int out_w,
                     const void* out, const void* aux2, const float* grad_out, float* grad_logits2,
                     int grad_is_zeroed, msq_stream_t stream);

/* ---------------------------------------------------------------------------
 * Source-side step, fused from LOW-resolution head logits: nn.CrossEntropyLoss(ignore_index=-1)
 * (tools/train_source.py:128,254,257) on the bilinearly upsampled logits, plus -- when cm != NULL --
 * np.argmax + Eval.add_batch of the same tensors (tools/train_source.py:280-283, utils/eval.py:109-121).
 *   label      int64 [N,out_h,out_w]; values outside [0,C) are ignored (the datasets emit -1)
 *   out.loss2  the mean over the out.nvalid_out counted pixels (NaN if none); out.ce_out their sum
 *   aux        float4-per-pixel cache (msq_fused_aux_bytes, nullable) consumed by msq_guidance_bwd,
 *              which is this loss's backward: dL/dlogits = grad/n_valid * (softmax - onehot(label))
 *   zero_grad  optional dL/dlogits buffer to zero-fill on the side
 *   cm         optional uint64 [C*C] confusion matrix, ACCUMULATED (row = label, column = argmax)
 * ------------------------------------------------------------------------- */
int msq_source_ce_fwd(const float* logits, const int64_t* label, int n, int num_class, int h, int w,
                      int out_h, int out_w, void* accum, void* out, void* aux /* nullable */,
                      float* zero_grad /* nullable */, unsigned long long* cm /* nullable */, msq_stream_t stream);

/* ---------------------------------------------------------------------------
 * Evaluation: Eval.__generate_matrix / add_batch (utils/eval.py:109-121).
 *   cm      uint64 [C*C], row = ground truth, column = prediction, ACCUMULATED
 *   errs    uint32 [2], OR-ed: [0] a flattened index C*gt+pred was negative
 *           (numpy.bincount raises ValueError), [1] it was >= C*C (the
 *           reference's reshape raises ValueError).  May be NULL for the
 *           logits variant (argmax is always in range).
 * gt values outside [0,C) are ignored (utils/eval.py:111).
 * ------------------------------------------------------------------------- */
int msq_confusion_i64(const int64_t* gt, const int64_t* pred, int64_t npix, int num_class,
                      unsigned long long* cm, unsigned int* errs, msq_stream_t stream);

/* Same with the callers' np.argmax(pred, axis=1) (tools/train_source.py:282,459)
 * fused in: logits float32 [N,C,H*W], first maximum wins, NaN counts as maximum. */
int msq_confusion_logits_f32(const int64_t* gt, const float* logits, int n, int num_class,
                             int64_t hw, unsigned long long* cm, msq_stream_t stream);

/* K (gt, pred) pairs in ONE launch: the reference's evaluation loops call Eval.add_batch once per image
 * (tools/train_source.py:429-492, tools/evaluate.py:99-202, tools/analysis.py:200-229), and one 8 MB launch per image is
 * launch-bound.  gt / pred / npix are HOST arrays of k device pointers / pixel counts (they travel as kernel parameters,
 * 32 pairs per launch).
 *   cm_stride = 0          every pair is ACCUMULATED into cm[C*C]   (k deferred add_batch calls)
 *   cm_stride >= C*C       pair j is ACCUMULATED into cm + j*cm_stride: one matrix per image, what tools/analysis.py:200-229
 *                          obtains with add_batch / metrics / reset() per image (no empty pairs in this mode)
 *   total (nullable)       every pair is additionally accumulated here (analysis.py's totalEval) */
int msq_confusion_i64_multi(const int64_t* const* gt, const int64_t* const* pred, const int64_t* npix, int k,
                            int num_class, unsigned long long* cm, int64_t cm_stride, unsigned long long* total,
                            unsigned int* errs, msq_stream_t stream);

/* Per-image matrices from a batch of fp32 logits [N,C,H*W] (argmax fused): image i is ACCUMULATED into
 * cm_per_image + i*C*C, and every image into total when it is not NULL. */
int msq_confusion_per_image_logits_f32(const int64_t* gt, const float* logits, int n, int num_class, int64_t hw,
                                       unsigned long long* cm_per_image, unsigned long long* total /* nullable */,
                                       msq_stream_t stream);

/* Flip-ensemble evaluation (tools/evaluate.py:120-141, --flip), fused: argmax_c of
 * (softmax(logits)[..., x] + softmax(logits_flipped)[..., W-1-x]) / 2 against gt, ACCUMULATED into cm.
 * logits_flipped is the model's output for the horizontally flipped image, NOT flipped back. */
int msq_confusion_flip_f32(const int64_t* gt, const float* logits, const float* logits_flipped, int n, int num_class,
                           int out_h, int out_w, unsigned long long* cm, msq_stream_t stream);

/* ---------------------------------------------------------------------------
 * Host-buffer pipeline (the one part of the library that owns memory): the fused
 * step for callers whose tensors live in HOST memory.  Each submission copies the
 * head logits H2D, runs msq_fused_fwd (+ msq_fused_bwd when host_grad != NULL) and
 * copies loss / histogram / dL/dlogits back, on one of `depth` internal streams,
 * so the copies of one step overlap the kernels of another.  The sequence mirrors
 * tools/solve_gta5.py:366-371,199,217 (x.to(device) ... loss ... backward ...
 * .item()).  Pin the host buffers for the copies to be asynchronous.
 *   grad_scale  the upstream gradient (lambda_target), passed by value
 *   slot_out    slot to hand to msq_pipe_wait before reading outputs / reusing inputs
 * ------------------------------------------------------------------------- */
typedef struct msq_pipe msq_pipe;
int  msq_pipe_create(int mode, int n, int num_class, int h, int w, int out_h, int out_w,
                     double ratio, int depth, msq_pipe** out);
int  msq_pipe_submit(msq_pipe* pipe, const float* host_logits, float grad_scale, float* host_loss,
                     float* host_grad /* nullable */, int32_t* host_hist /* nullable */, int* slot_out);
/* Images sharded over ranks: the steps use n_images_norm (the GLOBAL batch size) as the loss normaliser and exchange
 * their [loss | class histogram] vector through comm (msq_comm_*, below; NULL: no exchange), exactly as
 * msq_fused_fwd_bwd does.  Call once, before the first submit, on every rank. */
struct msq_comm;
int  msq_pipe_shard(msq_pipe* pipe, int n_images_norm, struct msq_comm* comm);
int  msq_pipe_wait(msq_pipe* pipe, int slot);
int  msq_pipe_drain(msq_pipe* pipe);
void msq_pipe_destroy(msq_pipe* pipe);

/* Performance-tuning knobs for bench sweeps ("conf_agg" 0|1|2, "conf_ctas" 1|2, "conf_grid" G,
 * "prob_waves" W, "fused_rows" R, "reserve_sms" S; 0 = automatic; "late_finalize" 0|1 and "pdl_mask" M: see api.cu).
 * Results never depend on them. */
int msq_tune_set(const char* key, int value);

/* ---------------------------------------------------------------------------
 * The one exchange of the image-sharded path: all-reduce(sum) of the packed fp64
 * statistics vector [loss | class hist(C) | confusion(C*C)] over NCCL / NVLink, enqueued by this
 * library on a communicator of its own (no torch.distributed call per step: ProcessGroupNCCL costs
 * ~25 us of host time per collective, most of a 35 us step).  Rank 0 calls msq_comm_unique_id and
 * hands the 128 bytes to the other ranks (any transport; the Python side uses a torch.distributed
 * broadcast); every rank then calls msq_comm_create (collective).  msq_comm_allreduce_f64 forks a
 * side stream from `stream`, so the collective overlaps whatever the caller enqueues next (the
 * backward kernel); msq_comm_join makes `stream` wait for the collective issued `lag` calls ago.  libnccl.so.2 is dlopen'ed.
 * ------------------------------------------------------------------------- */
typedef struct msq_comm msq_comm;
int msq_comm_unique_id(void* id128 /* host, 128 bytes */);
int msq_comm_create(const void* id128, int world, int rank, msq_comm** out);
int msq_comm_allreduce_f64(msq_comm* comm, double* buf /* device, in place */, int count, msq_stream_t stream);
/* the same for the path's INTEGER results (confusion counts of Eval, utils/eval.py:121; the adjacent
 * [ce_fix_out | nvalid_out] pair of the cross-entropy rows): uint64 words, exact whatever the order of the sum */
int msq_comm_allreduce_u64(msq_comm* comm, unsigned long long* buf /* device, in place */, int count, msq_stream_t stream);
/* The same sum for <= 2 words that the SAME step consumes (the [ce_fix_out | nvalid_out] pair: msq_guidance_bwd divides by the
 * global count): _begin right after the forward, independent work on the stream, _end right before the consumer.  Mailboxes
 * open: two 32-thread kernels, the words cross NVLink as 16-byte stores and a spinning warp sums them as integers (a few us
 * instead of a ~25 us collective launch); else one ncclAllReduce(uint64) on the side stream joined by _end.  One exchange in
 * flight per communicator; dst may equal src; every rank calls both, in the same order as its other mailbox steps. */
int msq_comm_sum_u64_begin(msq_comm* comm, const unsigned long long* src /* device */, int count, msq_stream_t stream);
int msq_comm_sum_u64_end(msq_comm* comm, unsigned long long* dst /* device */, int count, msq_stream_t stream);
int msq_comm_join(msq_comm* comm, int lag /* 0 = most recent all-reduce, k = k calls earlier (< 8) */, msq_stream_t stream);
void msq_comm_destroy(msq_comm* comm);

/* Peer-memory mailboxes (GPUs of one NVLink/NVSwitch box, <= 8 ranks): with them msq_fused_fwd_bwd does not call NCCL at
 * all.  An extra CTA of the step's backward kernel PUSHES the [loss | hist] vector this rank produced in the previous
 * step into every rank's mailbox with 16-byte {data, flag} stores over NVLink, and sums the vectors all ranks pushed for
 * the step before that, in rank order (bit-identical on every rank): no extra launch, no stream operation between the
 * step's kernels, no host cost, nothing on the forward -> backward critical path.
 * Ownership: the step's own vector and the all-reduced vectors live in device rings the COMMUNICATOR allocates (8 steps
 * deep); the caller's `out` keeps the rank-LOCAL statistics and is never read or written after the call that was given it
 * (it may be reused or freed at once, in stream order).  The all-reduced vector of step i is fetched with msq_comm_result
 * once step i+2 has been enqueued, or after msq_comm_join(comm, 0, stream), which completes the steps still in flight
 * (collective: every rank calls it).
 * Set-up (collective): every rank calls msq_comm_box_export (allocates its mailbox, returns a 64-byte cudaIpc handle), the
 * caller all-gathers the handles, every rank calls msq_comm_box_open with all of them in rank order (maps the peers), the
 * ranks agree on whether ALL of them succeeded, and every rank calls msq_comm_box_enable(comm, 1) -- or none does, and the
 * communicator keeps using ncclAllReduce (CUDA IPC not permitted, no peer access, more than 8 ranks).
 * Liveness: a reduction waits for a peer's vector for msq_comm_box_timeout seconds (default 600, or MSQ_BOX_TIMEOUT_S);
 * after that the vector of THAT step becomes NaN, the error is returned as MSQ_E_PEER by the next msq_fused_fwd_bwd /
 * msq_comm_join (once per loss), and the next step waits afresh: a slow rank costs the statistics it missed, nothing else. */
int msq_comm_box_export(msq_comm* comm, void* handle64 /* host, 64 bytes out */);
int msq_comm_box_open(msq_comm* comm, const void* handles /* host, world x 64 bytes, rank order */);
int msq_comm_box_enable(msq_comm* comm, int on);               /* on: needs a successful box_open; no steps in flight */
int msq_comm_box_active(const msq_comm* comm);                 /* 1: mailboxes in use, 0: NCCL */
int msq_comm_box_errors(msq_comm* comm, unsigned* out);        /* bit 0: a peer's vector did not arrive in time (synchronises) */
int msq_comm_box_timeout(msq_comm* comm, double seconds);      /* per-vector wait limit (synchronises the device) */

/* All-reduced [loss | class hist] of the msq_fused_fwd_bwd step issued `lag` steps before the most recent one (0 = the most
 * recent), copied device-to-device into dst (count <= 1 + C doubles) on `stream`.  MSQ_E_NOTREADY if that vector has not
 * been produced yet (mailboxes: two steps later or after msq_comm_join; at most 6 steps back). */
int msq_comm_result(msq_comm* comm, int lag, double* dst, int count, msq_stream_t stream);

/* One library call per training step, for callers that know the upstream gradient when they call the forward -- lambda_target
 * is a constant (tools/solve_gta5.py:199,217).  grad = *grad_out (device scalar) if grad_out != NULL, else grad_scale.
 * Same results as msq_fused_fwd followed by msq_fused_bwd -- every output bit-identical, the gradient up to the order of its
 * atomic adds -- from TWO kernels instead of three: the backward directly follows the forward, derives the image-wise weights
 * from the forward's replicated class histogram itself (the arithmetic of the finalisation), and carries the finalisation in
 * an extra CTA that runs beside its rows and zeroes the accumulators once every work CTA has taken what it needs (the
 * `ticket` word of `accum` counts them; it is zero again afterwards).  The finalisation kernel's launch, two grid-completion
 * hand-overs and its load / powf / fp64 latency leave the step's critical path: 30.9 -> 28.6 us at 2 x 512 x 1024 pixels on
 * B200.  (Tuning knob "late_finalize" = 0 restores forward -> finalisation -> backward.)
 * comm != NULL: the step's statistics vector is all-reduced as well -- pushed / reduced over the peer-memory mailboxes by a
 * second extra CTA of the backward when they are open (result two steps later, see above), else an ncclAllReduce of the
 * communicator's copy forked after the backward and ordered after the collective issued `lag` steps earlier.
 * out.stats holds the rank-LOCAL vector; fetch the all-reduced one with msq_comm_result.
 * Ordering contract of every fused kernel chain in this library: the kernels are launched with programmatic stream
 * serialisation, and the BACKWARD kernels stage their logits tile before waiting for the preceding kernel, so `logits` must be
 * complete when the step's first kernel is launched; a producer that itself triggers dependents early
 * (griddepcontrol.launch_dependents before its last store) must not directly precede them in the stream.  The forward kernels
 * read nothing before their wait. */
int msq_fused_fwd_bwd(int mode, const float* logits, int n, int num_class, int h, int w, int out_h, int out_w,
                      double ratio, int n_images_norm, void* accum, void* out, void* aux /* nullable */,
                      const float* grad_out /* nullable */, float grad_scale, float* grad_logits,
                      msq_comm* comm /* nullable */, int lag, msq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MSQ_B200_H_ */
