// Stock layers that sit between the MNF layers in the reference models
// (mnf_lenet.py:23-32: ReLU, MaxPool2D(padding="SAME"), Flatten, Softmax;
//  mnf_feed_forward.py:20: ReLU, inference-mode BatchNormalization) plus the training-step
// closure pieces (lenet_mnist.py:53-57 cross-entropy, :47,115 Adam).  SURVEY.md 8(f)-1,3.
#include "common.cuh"

static inline int ew_grid(mnf_ctx* ctx, int64_t n, int threads = 256) {
  int64_t g = cdiv(n, threads), cap = (int64_t)ctx->sm_count * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ReLU then 2x2/stride-2 max-pool with TF "SAME" padding (pad only at bottom/right for odd sizes)
__global__ void relu_maxpool2_kernel(int64_t B, int H, int W, int C, int Ho, int Wo, const float* __restrict__ x,
                                     float* __restrict__ y, int8_t* __restrict__ argmax) {
  const int64_t total = B * Ho * Wo * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int wo = (int)(t % Wo);
    t /= Wo;
    int ho = (int)(t % Ho);
    int64_t b = t / Ho;
    float best = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int hi = 2 * ho + (q >> 1), wi = 2 * wo + (q & 1);
      if (hi < H && wi < W) {
        float v = __ldg(x + ((b * H + hi) * W + wi) * C + c);
        if (v > best) { best = v; arg = q; }
      }
    }
    y[i] = fmaxf(best, 0.f);
    if (argmax) argmax[i] = (int8_t)arg;
  }
}

__global__ void relu_maxpool2_bwd_kernel(int64_t B, int H, int W, int C, int Ho, int Wo, const float* __restrict__ x,
                                         const int8_t* __restrict__ argmax, const float* __restrict__ dy,
                                         float* __restrict__ dx) {
  const int64_t total = B * H * W * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int wi = (int)(t % W);
    t /= W;
    int hi = (int)(t % H);
    int64_t b = t / H;
    int ho = hi >> 1, wo = wi >> 1;
    int q = ((hi & 1) << 1) | (wi & 1);
    int64_t o = ((b * Ho + ho) * Wo + wo) * C + c;
    float g = (argmax[o] == q && x[i] > 0.f) ? dy[o] : 0.f;
    dx[i] = g;
  }
}

__global__ void relu_kernel(int64_t n, const float* __restrict__ x, float* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fmaxf(x[i], 0.f);
}
__global__ void relu_bwd_kernel(int64_t n, const float* __restrict__ x, const float* __restrict__ dy,
                                float* __restrict__ dx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = x[i] > 0.f ? dy[i] : 0.f;
}
__global__ void scale_kernel(int64_t n, float a, const float* __restrict__ x, float* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a * x[i];
}
__global__ void axpy_kernel(int64_t n, float a, const float* __restrict__ x, float* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fmaf(a, x[i], y[i]);
}

// warp per row softmax (N small: class count)
__global__ void softmax_kernel(int64_t B, int N, const float* __restrict__ x, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) mx = fmaxf(mx, x[row * N + j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int j = lane; j < N; j += 32) s += expf(x[row * N + j] - mx);
  s = warp_sum(s);
  for (int j = lane; j < N; j += 32) y[row * N + j] = expf(x[row * N + j] - mx) / s;
}

// loss = mean_b( -sum_j labels * log softmax ), d_logits = (softmax * sum(labels) - labels) * inv_global_batch
__global__ void softmax_xent_kernel(int64_t B, int N, const float* __restrict__ x, const float* __restrict__ lab,
                                    float inv_gb, float* loss_out, float* __restrict__ dlog) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  float l = 0.f;
  if (row < B) {
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, x[row * N + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f, ls = 0.f;
    for (int j = lane; j < N; j += 32) {
      s += expf(x[row * N + j] - mx);
      ls += lab[row * N + j];
    }
    s = warp_sum(s);
    ls = warp_sum(ls);
    float lse = mx + logf(s);
    for (int j = lane; j < N; j += 32) {
      float lp = x[row * N + j] - lse;
      l -= lab[row * N + j] * lp;
      if (dlog) dlog[row * N + j] = (expf(lp) * ls - lab[row * N + j]) * inv_gb;
    }
    l = warp_sum(l);
  }
  if (lane == 0 && row < B) atomicAdd(loss_out, l * inv_gb);
}

// loss = mean((y - t)^2) over all n elements of the GLOBAL batch (bandgap_regression.py:57-69);
// dy = 2 (y - t) * inv_n
__global__ void mse_kernel(int64_t n, const float* __restrict__ y, const float* __restrict__ t, float inv_n,
                           float* loss_out, float* __restrict__ dy) {
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = y[i] - t[i];
    acc = fmaf(d, d, acc);
    if (dy) dy[i] = 2.f * d * inv_n;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss_out, acc * inv_n);
}

__global__ void adam_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, float lr_t, float b1, float b2, float eps) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}


// Adam with the step count on the DEVICE (a captured training step replays with a fresh bias correction):
// state = {int32 step, float lr_t}; adam_dev_prepare bumps the step and derives lr_t as mnf_adam_step does.
__global__ void adam_dev_prepare(int32_t* state, float lr, float b1, float b2) {
  const int t = ++state[0];
  reinterpret_cast<float*>(state)[1] = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
}
__global__ void adam_dev_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, const int32_t* __restrict__ state, float b1, float b2, float eps) {
  const float lr_t = reinterpret_cast<const float*>(state)[1];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// ---- tf.keras.layers.BatchNormalization over the last axis of [B, N] (mnf_feed_forward.py:20; Keras
// defaults momentum 0.99, epsilon 1e-3).  Block = 32 columns x 8 row groups: lanes walk columns
// (coalesced), row groups stride the batch; reductions over the batch through shared memory,
// accumulated in double (B terms of either sign).
//   training: batch mean and BIASED variance normalise (the rank-2, non-fused Keras path), and the
//             moving statistics move by (1 - momentum) towards them;
//   inference: the moving statistics normalise.
constexpr int BN_COLS = 32, BN_RG = 8;

__device__ __forceinline__ double bn_block_sum(double v, double (*red)[BN_COLS]) {
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  red[rg][c] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int g = 0; g < BN_RG; ++g) s += red[g][c];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(BN_COLS* BN_RG) batchnorm_fwd_kernel(
    int64_t B, int N, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ mov_mean, float* __restrict__ mov_var, float eps, float momentum, int training,
    float* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  __shared__ double red[BN_RG][BN_COLS];
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int n = blockIdx.x * BN_COLS + c;
  const bool ok = n < N;
  float mean, var;
  if (training) {
    double s = 0.0;
    if (ok)
      for (int64_t b = rg; b < B; b += BN_RG) s += (double)x[b * N + n];
    const double m = bn_block_sum(s, red) / (double)B;
    double q = 0.0;
    if (ok)
      for (int64_t b = rg; b < B; b += BN_RG) {
        const double d = (double)x[b * N + n] - m;
        q += d * d;
      }
    const double v = bn_block_sum(q, red) / (double)B;
    mean = (float)m;
    var = (float)v;
    if (ok && rg == 0) {
      mov_mean[n] = mov_mean[n] * momentum + mean * (1.f - momentum);
      mov_var[n] = mov_var[n] * momentum + var * (1.f - momentum);
    }
  } else {
    mean = ok ? mov_mean[n] : 0.f;
    var = ok ? mov_var[n] : 1.f;
  }
  if (!ok) return;
  const float inv = 1.0f / sqrtf(var + eps);
  if (rg == 0 && save_mean) { save_mean[n] = mean; save_invstd[n] = inv; }
  const float g = gamma[n], bt = beta[n];
  for (int64_t b = rg; b < B; b += BN_RG) y[b * N + n] = (x[b * N + n] - mean) * inv * g + bt;
}

// training: dx = gamma * inv / B * (B dy - sum(dy) - xhat sum(dy xhat));  inference: dx = dy gamma inv
__global__ void __launch_bounds__(BN_COLS* BN_RG) batchnorm_bwd_kernel(
    int64_t B, int N, const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ save_mean,
    const float* __restrict__ save_invstd, int training, const float* __restrict__ dy, float* __restrict__ dx,
    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double red[BN_RG][BN_COLS];
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int n = blockIdx.x * BN_COLS + c;
  const bool ok = n < N;
  const float mean = ok ? save_mean[n] : 0.f, inv = ok ? save_invstd[n] : 0.f;
  double s1 = 0.0, s2 = 0.0;
  if (ok)
    for (int64_t b = rg; b < B; b += BN_RG) {
      const float g = dy[b * N + n];
      s1 += (double)g;
      s2 += (double)g * (double)((x[b * N + n] - mean) * inv);
    }
  const double sum_dy = bn_block_sum(s1, red), sum_dyx = bn_block_sum(s2, red);
  if (!ok) return;
  if (rg == 0) {
    if (dgamma) dgamma[n] = (float)sum_dyx;
    if (dbeta) dbeta[n] = (float)sum_dy;
  }
  if (!dx) return;
  const float gi = gamma[n] * inv;
  const float m1 = (float)(sum_dy / (double)B), m2 = (float)(sum_dyx / (double)B);
  for (int64_t b = rg; b < B; b += BN_RG) {
    const float g = dy[b * N + n];
    dx[b * N + n] = training ? gi * (g - m1 - (x[b * N + n] - mean) * inv * m2) : gi * g;
  }
}

extern "C" {

int mnf_relu_maxpool2(mnf_ctx* ctx, int64_t B, int64_t H, int64_t W, int64_t C, const float* x, float* y,
                      int8_t* argmax) {
  MNF_CHECK_ARG(ctx && x && y && B >= 0 && H >= 1 && W >= 1 && C >= 1, "mnf_relu_maxpool2: bad argument");
  if (B == 0) return MNF_OK;
  int Ho = (int)cdiv(H, 2), Wo = (int)cdiv(W, 2);
  relu_maxpool2_kernel<<<ew_grid(ctx, B * Ho * Wo * C), 256, 0, ctx->stream>>>(B, (int)H, (int)W, (int)C, Ho, Wo, x, y,
                                                                            argmax);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_relu_maxpool2_backward(mnf_ctx* ctx, int64_t B, int64_t H, int64_t W, int64_t C, const float* x,
                               const int8_t* argmax, const float* dy, float* dx) {
  MNF_CHECK_ARG(ctx && x && argmax && dy && dx && B >= 0 && H >= 1 && W >= 1 && C >= 1,
                "mnf_relu_maxpool2_backward: bad argument");
  if (B == 0) return MNF_OK;
  int Ho = (int)cdiv(H, 2), Wo = (int)cdiv(W, 2);
  relu_maxpool2_bwd_kernel<<<ew_grid(ctx, B * H * W * C), 256, 0, ctx->stream>>>(B, (int)H, (int)W, (int)C, Ho, Wo, x,
                                                                              argmax, dy, dx);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_relu(mnf_ctx* ctx, int64_t n, const float* x, float* y) {
  MNF_CHECK_ARG(ctx && (n == 0 || (x && y)) && n >= 0, "mnf_relu: bad argument");
  if (n == 0) return MNF_OK;
  relu_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, x, y);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_relu_backward(mnf_ctx* ctx, int64_t n, const float* x, const float* dy, float* dx) {
  MNF_CHECK_ARG(ctx && (n == 0 || (x && dy && dx)) && n >= 0, "mnf_relu_backward: bad argument");
  if (n == 0) return MNF_OK;
  relu_bwd_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, x, dy, dx);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_softmax(mnf_ctx* ctx, int64_t B, int64_t N, const float* x, float* y) {
  MNF_CHECK_ARG(ctx && x && y && B >= 0 && N >= 1 && N < (1 << 30), "mnf_softmax: bad argument");
  if (B == 0) return MNF_OK;
  softmax_kernel<<<(unsigned)cdiv(B, 8), 256, 0, ctx->stream>>>(B, (int)N, x, y);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_softmax_xent(mnf_ctx* ctx, int64_t B, int64_t N, const float* logits, const float* labels,
                     float inv_global_batch, float* loss_out, float* d_logits) {
  MNF_CHECK_ARG(ctx && logits && labels && loss_out && B >= 0 && N >= 1 && N < (1 << 30), "mnf_softmax_xent: bad argument");
  MNF_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), ctx->stream));
  if (B == 0) return MNF_OK;
  softmax_xent_kernel<<<(unsigned)cdiv(B, 8), 256, 0, ctx->stream>>>(B, (int)N, logits, labels, inv_global_batch,
                                                                     loss_out, d_logits);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_mse_loss(mnf_ctx* ctx, int64_t n, const float* y, const float* target, float inv_global_n, float* loss_out,
                 float* dy) {
  MNF_CHECK_ARG(ctx && y && target && loss_out && n >= 0, "mnf_mse_loss: bad argument");
  MNF_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), ctx->stream));
  if (n == 0) return MNF_OK;
  mse_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, y, target, inv_global_n, loss_out, dy);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_batchnorm_forward(mnf_ctx* ctx, int64_t B, int64_t N, const float* x, const float* gamma, const float* beta,
                          float* moving_mean, float* moving_var, float eps, float momentum, int32_t training, float* y,
                          float* save_mean, float* save_invstd) {
  MNF_CHECK_ARG(ctx && x && gamma && beta && moving_mean && moving_var && y, "mnf_batchnorm_forward: NULL argument");
  MNF_CHECK_ARG(B >= 1 && N >= 1 && N < (1 << 30), "mnf_batchnorm_forward: bad shape [%lld,%lld]", (long long)B, (long long)N);
  MNF_CHECK_ARG((save_mean == nullptr) == (save_invstd == nullptr), "mnf_batchnorm_forward: save_mean / save_invstd go together");
  batchnorm_fwd_kernel<<<(int)cdiv(N, BN_COLS), BN_COLS * BN_RG, 0, ctx->stream>>>(
      B, (int)N, x, gamma, beta, moving_mean, moving_var, eps, momentum, training, y, save_mean, save_invstd);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_batchnorm_backward(mnf_ctx* ctx, int64_t B, int64_t N, const float* x, const float* gamma, const float* save_mean,
                           const float* save_invstd, int32_t training, const float* dy, float* dx, float* d_gamma,
                           float* d_beta) {
  MNF_CHECK_ARG(ctx && x && gamma && save_mean && save_invstd && dy, "mnf_batchnorm_backward: NULL argument");
  MNF_CHECK_ARG(B >= 1 && N >= 1 && N < (1 << 30), "mnf_batchnorm_backward: bad shape [%lld,%lld]", (long long)B, (long long)N);
  batchnorm_bwd_kernel<<<(int)cdiv(N, BN_COLS), BN_COLS * BN_RG, 0, ctx->stream>>>(B, (int)N, x, gamma, save_mean, save_invstd,
                                                                                 training, dy, dx, d_gamma, d_beta);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_scale(mnf_ctx* ctx, int64_t n, float a, const float* x, float* y) {
  MNF_CHECK_ARG(ctx && (n == 0 || (x && y)) && n >= 0, "mnf_scale: bad argument");
  if (n == 0) return MNF_OK;
  scale_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, a, x, y);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

__global__ void tile_rows_kernel(int64_t total4, int64_t row4, int64_t repeat, const float4* __restrict__ src,
                                 float4* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t orow = i / row4, e = i - orow * row4;
    dst[i] = __ldg(src + (orow / repeat) * row4 + e);
  }
}
__global__ void tile_rows_kernel1(int64_t total, int64_t row, int64_t repeat, const float* __restrict__ src,
                                  float* __restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t orow = i / row, e = i - orow * row;
    dst[i] = __ldg(src + (orow / repeat) * row + e);
  }
}

__global__ void u32_add_kernel(uint32_t* p, uint32_t inc) { *p += inc; }

int mnf_u32_add(mnf_ctx* ctx, uint32_t* counter, uint32_t inc) {
  MNF_CHECK_ARG(ctx && counter, "mnf_u32_add: NULL argument");
  u32_add_kernel<<<1, 1, 0, ctx->stream>>>(counter, inc);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_tile_rows(mnf_ctx* ctx, int64_t n_rows, int64_t row_elems, int64_t repeat, const float* src, float* dst) {
  MNF_CHECK_ARG(ctx && n_rows >= 0 && row_elems >= 0 && repeat >= 1, "mnf_tile_rows: bad argument");
  const int64_t total = n_rows * repeat * row_elems;
  if (total == 0) return MNF_OK;
  MNF_CHECK_ARG(src && dst, "mnf_tile_rows: null buffer");
  if ((row_elems & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0)
    tile_rows_kernel<<<ew_grid(ctx, total / 4), 256, 0, ctx->stream>>>(total / 4, row_elems / 4, repeat,
                                                                         (const float4*)src, (float4*)dst);
  else
    tile_rows_kernel1<<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(total, row_elems, repeat, src, dst);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_axpy(mnf_ctx* ctx, int64_t n, float a, const float* x, float* y) {
  MNF_CHECK_ARG(ctx && (n == 0 || (x && y)) && n >= 0, "mnf_axpy: bad argument");
  if (n == 0) return MNF_OK;
  axpy_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, a, x, y);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_adam_step(mnf_ctx* ctx, int64_t n, float* param, const float* grad, float* m, float* v, float lr, float beta1,
                  float beta2, float eps, int32_t step) {
  MNF_CHECK_ARG(ctx && (n == 0 || (param && grad && m && v)) && n >= 0 && step >= 1, "mnf_adam_step: bad argument");
  if (n == 0) return MNF_OK;
  // Keras/TF Adam: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); update = lr_t * m / (sqrt(v) + eps)
  double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  adam_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, param, grad, m, v, (float)lr_t, beta1, beta2, eps);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_adam_step_dev(mnf_ctx* ctx, int64_t n, float* param, const float* grad, float* m, float* v, float lr, float beta1,
                      float beta2, float eps, int32_t* d_state) {
  MNF_CHECK_ARG(ctx && d_state && (n == 0 || (param && grad && m && v)) && n >= 0, "mnf_adam_step_dev: bad argument");
  adam_dev_prepare<<<1, 1, 0, ctx->stream>>>(d_state, lr, beta1, beta2);
  MNF_LAUNCHED(ctx);
  if (n == 0) return MNF_OK;
  adam_dev_kernel<<<ew_grid(ctx, n), 256, 0, ctx->stream>>>(n, param, grad, m, v, d_state, beta1, beta2, eps);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

}  // extern "C"
