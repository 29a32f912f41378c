// Backward entry points (TF autodiff of the hot path).  Placeholder
// bodies: they fail loudly until the kernels land (never a silent fallback).
#include "common.cuh"

#define NOT_YET(name)                                                   \
  do {                                                                  \
    mnf_set_error(name ": backward kernel not implemented yet");        \
    return MNF_ERR_UNSUPPORTED;                                         \
  } while (0)

extern "C" {
int mnf_flow_backward(mnf_ctx*, int64_t, int64_t, int32_t, const mnf_flow_step*, const float*, const float*,
                      const float*, const float*, float*, const mnf_flow_step*) { NOT_YET("mnf_flow_backward"); }
int mnf_sample_z_backward(mnf_ctx*, int64_t, int64_t, const float*, const float*, const mnf_noise*, float*, float*) {
  NOT_YET("mnf_sample_z_backward");
}
int mnf_dense_backward(mnf_ctx*, int64_t, int64_t, int64_t, const float*, const float*, const float*, const float*,
                       const float*, float, const mnf_noise*, const float*, const float*, float*, float*, float*,
                       float*, float*, float*) { NOT_YET("mnf_dense_backward"); }
int mnf_conv2d_backward(mnf_ctx*, const mnf_conv_shape*, const float*, const float*, const float*, const float*,
                        const float*, float, const mnf_noise*, const float*, const float*, const float*, float*,
                        float*, float*, float*, float*, float*) { NOT_YET("mnf_conv2d_backward"); }
int mnf_kl_backward(mnf_ctx*, const mnf_kl_args*, float, mnf_kl_grads*) { NOT_YET("mnf_kl_backward"); }
}
