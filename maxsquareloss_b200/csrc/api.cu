// ABI housekeeping: version, error names, state layout, tuning knobs.
#include <string.h>
#include "common.cuh"

namespace msq {
extern int g_conf_agg;      // confusion.cu
extern int g_fused_rows;    // fused_loss.cu
}

extern "C" int msq_abi_version(void) { return MSQ_ABI_VERSION; }

extern "C" const char* msq_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case MSQ_E_BADARG: return "msq: bad argument (null pointer, non-positive size or more than 32 classes)";
        case MSQ_E_GEOMETRY: return "msq: fused path needs out_h >= h and out_w >= w (bilinear upsampling only)";
        case MSQ_E_SMEM: return "msq: low-resolution tile does not fit in shared memory";
        case MSQ_E_ALIGN: return "msq: pointer is not aligned for its element type";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "msq: unknown error";
    }
}

extern "C" int msq_state_layout_get(int n_images, int num_class, msq_state_layout* out) {
    if (!out || n_images < 1 || num_class < 1 || num_class > MSQ_MAX_CLASSES) return MSQ_E_BADARG;
    *out = msq::make_layout(n_images, num_class);
    return 0;
}

// Performance-tuning knobs (bench sweeps); results never depend on them.
//   "conf_agg"   0|1|2  warp aggregation level of the confusion histogram
//   "fused_rows" R      output rows per strip of the fused kernels (0 = automatic)
extern "C" int msq_tune_set(const char* key, int value) {
    if (!key) return MSQ_E_BADARG;
    if (!strcmp(key, "conf_agg")) { msq::g_conf_agg = value; return 0; }
    if (!strcmp(key, "fused_rows")) { msq::g_fused_rows = value; return 0; }
    return MSQ_E_BADARG;
}
