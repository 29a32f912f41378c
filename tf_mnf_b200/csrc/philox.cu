// Materialise the counter-based noise the fused kernels draw in registers, so that tests
// (and users who need exact reproducibility against another implementation) can read it.
// Replaces tf.random.normal (mnf_dense.py:94,141,167; mnf_conv.py:106,156,165,194) and
// keras.backend.random_binomial (rnvp.py:36).
#include "common.cuh"

__global__ void philox_normal_kernel(NoiseDev n, int64_t rows, int64_t outer, int64_t inner, float* out) {
  const int64_t inner4 = (inner + 3) >> 2;
  const int64_t total = rows * outer * inner4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t q = i % inner4;
    int64_t o = (i / inner4) % outer;
    int64_t r = i / (inner4 * outer);
    float v[4];
    philox_normal4(n, n.row_offset + (uint64_t)r, (uint32_t)o, (uint32_t)q, v);
    float* dst = out + (r * outer + o) * inner + q * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (q * 4 + k < inner) dst[k] = v[k];
  }
}

__global__ void philox_raw_kernel(NoiseDev n, int64_t rows, int64_t outer, int64_t inner, uint32_t* out) {
  const int64_t inner4 = (inner + 3) >> 2;
  const int64_t total = rows * outer * inner4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t q = i % inner4;
    int64_t o = (i / inner4) % outer;
    int64_t r = i / (inner4 * outer);
    Philox4 p = philox_at(n, n.row_offset + (uint64_t)r, (uint32_t)o, (uint32_t)q);
    uint32_t v[4] = {p.x, p.y, p.z, p.w};
    uint32_t* dst = out + (r * outer + o) * inner + q * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (q * 4 + k < inner) dst[k] = v[k];
  }
}

__global__ void philox_bern_kernel(NoiseDev n, int64_t rows, int64_t inner, float p, float* out) {
  const int64_t total = rows * inner;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / inner, d = i % inner;
    out[i] = noise_bern1(n, r, (uint32_t)d, inner, p);
  }
}

static int grid_for(mnf_ctx* ctx, int64_t total) {
  int64_t g = cdiv(total, 256);
  int64_t cap = (int64_t)ctx->sm_count * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" {

int mnf_philox_normal(mnf_ctx* ctx, uint64_t seed, uint32_t stream, uint64_t row_offset, int64_t rows,
                      int64_t outer, int64_t inner, float* out) {
  MNF_CHECK_ARG(ctx && out, "mnf_philox_normal: NULL argument");
  MNF_CHECK_ARG(rows >= 0 && outer >= 1 && inner >= 1, "mnf_philox_normal: bad shape");
  if (rows == 0) return MNF_OK;
  mnf_noise nz = {seed, row_offset, stream, 0, nullptr};
  NoiseDev n = make_noise(&nz);
  int64_t total = rows * outer * ((inner + 3) / 4);
  philox_normal_kernel<<<grid_for(ctx, total), 256, 0, ctx->stream>>>(n, rows, outer, inner, out);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_philox_raw(mnf_ctx* ctx, uint64_t seed, uint32_t stream, uint64_t row_offset, int64_t rows,
                   int64_t outer, int64_t inner, uint32_t* out) {
  MNF_CHECK_ARG(ctx && out, "mnf_philox_raw: NULL argument");
  MNF_CHECK_ARG(rows >= 0 && outer >= 1 && inner >= 1, "mnf_philox_raw: bad shape");
  if (rows == 0) return MNF_OK;
  mnf_noise nz = {seed, row_offset, stream, 0, nullptr};
  NoiseDev n = make_noise(&nz);
  int64_t total = rows * outer * ((inner + 3) / 4);
  philox_raw_kernel<<<grid_for(ctx, total), 256, 0, ctx->stream>>>(n, rows, outer, inner, out);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_philox_bernoulli(mnf_ctx* ctx, uint64_t seed, uint32_t stream, uint64_t row_offset, int64_t rows,
                         int64_t inner, float p, float* out) {
  MNF_CHECK_ARG(ctx && out, "mnf_philox_bernoulli: NULL argument");
  MNF_CHECK_ARG(rows >= 0 && inner >= 1, "mnf_philox_bernoulli: bad shape");
  if (rows == 0) return MNF_OK;
  mnf_noise nz = {seed, row_offset, stream, 0, nullptr};
  NoiseDev n = make_noise(&nz);
  philox_bern_kernel<<<grid_for(ctx, rows * inner), 256, 0, ctx->stream>>>(n, rows, inner, p, out);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

}  // extern "C"
