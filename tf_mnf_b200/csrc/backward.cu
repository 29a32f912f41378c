// Backward of the MNF layer hot path: what TF autodiff produces for MNFDense.call /
// MNFConv2D.call / sample_z / the flow chains (the reference has no hand-written gradients;
// formulas: SURVEY.md App. A-1, A-1c, A-2, validated against finite differences in
// tests/test_oracle_grad.py).  Correctness-first fp32 kernels: a strided SGEMM with split-K,
// explicit im2col/col2im for the conv gradients, and fused row-tile kernels for the flows.
#include <cstdlib>

#include "common.cuh"

// ----------------------------------------------------------------------- workspace bump
struct Bump {
  char* base;
  size_t off;
  template <typename T>
  T* take(size_t n) {
    T* p = (T*)(base + off);
    off += (n * sizeof(T) + 255) & ~(size_t)255;
    return p;
  }
};
static inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

static inline int ew_grid(mnf_ctx* ctx, int64_t n) {
  int64_t g = cdiv(n, 256), cap = (int64_t)ctx->sm_count * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ------------------------------------------------------------------------ strided SGEMM
// C[m,n] (+)= sum_k A[m*sAm + k*sAk] * B[k*sBk + n*sBn];  64x64x16 tiles, 4x4 per thread.
// gridDim.z > 1 => split-K with atomicAdd (C must hold the value to accumulate onto).
struct GemmArgs {
  int M, N, K;
  const float* A; int64_t sAm, sAk;
  const float* B; int64_t sBk, sBn;
  float* C; int64_t ldc;
  int accumulate;
  int kchunk;  // K range per z-slice
  int sqA;     // use A(m,k)^2 (the variance twin of a contraction reads the same operand)
};

__global__ void __launch_bounds__(256) sgemm_strided(GemmArgs g) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int kbeg = blockIdx.z * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
  float acc[4][4] = {};
  const bool a_kfast = (g.sAk == 1), b_nfast = (g.sBn == 1);
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int m, k;
      if (a_kfast) { k = tid & 15; m = (tid >> 4) + 16 * q; } else { m = tid & 63; k = (tid >> 6) + 4 * q; }
      float v = 0.f;
      if (m0 + m < g.M && k0 + k < kend) v = __ldg(g.A + (int64_t)(m0 + m) * g.sAm + (int64_t)(k0 + k) * g.sAk);
      As[k][m] = g.sqA ? v * v : v;
      int n, kb;
      if (b_nfast) { n = tid & 63; kb = (tid >> 6) + 4 * q; } else { kb = tid & 15; n = (tid >> 4) + 16 * q; }
      float w = 0.f;
      if (n0 + n < g.N && k0 + kb < kend) w = __ldg(g.B + (int64_t)(k0 + kb) * g.sBk + (int64_t)(n0 + n) * g.sBn);
      Bs[kb][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float* c = g.C + (int64_t)m * g.ldc + n;
      if (gridDim.z > 1) atomicAdd(c, acc[i][j]);
      else *c = g.accumulate ? *c + acc[i][j] : acc[i][j];
    }
  }
}

static int gemm(mnf_ctx* ctx, int M, int N, int K, const float* A, int64_t sAm, int64_t sAk, const float* B,
                int64_t sBk, int64_t sBn, float* C, int64_t ldc, bool accumulate, bool sqA = false) {
  if (M <= 0 || N <= 0) return MNF_OK;
  // tensor-core math modes: 3xTF32 on tcgen05 (gemm_umma.cu) unless the product is tiny
  static const bool no_tc = getenv("MNF_NO_TC_BWD") != nullptr;
  if (ctx->math_mode != MNF_MATH_FP32 && !no_tc && K >= 1 && (int64_t)M * N * K >= (1 << 18))
    return mnf_gemm_umma(ctx, M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc, accumulate, sqA);
  GemmArgs g = {M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc, accumulate ? 1 : 0, K, sqA ? 1 : 0};
  dim3 grid((unsigned)cdiv(M, 64), (unsigned)cdiv(N, 64), 1);
  int tiles = grid.x * grid.y;
  int split = 1;
  if (tiles < ctx->sm_count && K >= 512 && (accumulate || ldc == N)) {
    split = (int)cdiv(2 * ctx->sm_count, tiles);
    int maxsplit = (int)cdiv(K, 128);
    if (split > maxsplit) split = maxsplit;
    if (split < 1) split = 1;
  }
  if (split > 1) {
    g.kchunk = (int)(cdiv(cdiv(K, split), 16) * 16);
    grid.z = (unsigned)cdiv(K, g.kchunk);
    if (!accumulate) {
      MNF_CHECK_ARG(ldc == N, "internal: split-K needs a compact C");
      MNF_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, ctx->stream));
    }
  }
  sgemm_strided<<<grid, 256, 0, ctx->stream>>>(g);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

// the same GEMM for other translation units (flows.cu: wide IAF steps on few rows)
int mnf_gemm_any(mnf_ctx* ctx, int M, int N, int K, const float* A, int64_t sAm, int64_t sAk, const float* B,
                 int64_t sBk, int64_t sBn, float* C, int64_t ldc, bool accumulate, bool sqA) {
  return gemm(ctx, M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc, accumulate, sqA);
}

// ------------------------------------------------------------------- shared small kernels
__global__ void bwd_prep_var(int64_t nW, const float* __restrict__ Ls, float max_std, float* __restrict__ Vw, int N,
                             const float* __restrict__ lvb, float* __restrict__ vb) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nW; i += (int64_t)gridDim.x * blockDim.x) {
    float s = fminf(expf(Ls[i]), max_std);
    Vw[i] = s * s;
    if (i < N) vb[i] = fminf(expf(lvb[i]), max_std * max_std);
  }
}

// dV = dy * eps / (2 sd); conv additionally dm = dy * z[b,f]   (rows = B for dense, B*Ho*Wo for conv)
__global__ void dy_split_kernel(int64_t M, int N, int64_t HoWo, const float* __restrict__ dy,
                                const float* __restrict__ sd, const float* __restrict__ z, NoiseDev eps,
                                float* __restrict__ dV, float* __restrict__ dm) {
  const int64_t total = M * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t m = i / N;
    int n = (int)(i - m * N);
    int64_t b = m / HoWo;
    uint32_t outer = (uint32_t)(m - b * HoWo);
    float e = eps.injected ? eps.injected[i] : noise_normal1(eps, b, outer, (uint32_t)n, HoWo, N);
    float g = dy[i];
    dV[i] = g * e / (2.f * sd[i]);
    if (dm) dm[i] = z ? g * z[b * N + n] : g;
  }
}

// N % 4 == 0: a thread per four consecutive columns -- one Philox call instead of four, 16-byte accesses
__global__ void dy_split_kernel4(int64_t M, int N, int64_t HoWo, const float* __restrict__ dy,
                                 const float* __restrict__ sd, const float* __restrict__ z, NoiseDev eps,
                                 float* __restrict__ dV, float* __restrict__ dm) {
  const int N4 = N >> 2;
  const int64_t total = M * N4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = q / N4;
    const int n = (int)(q - m * N4) * 4;
    const int64_t b = m / HoWo, i = m * N + n;
    const uint32_t outer = (uint32_t)(m - b * HoWo);
    float e[4];
    if (eps.injected) {
      const float4 t = *reinterpret_cast<const float4*>(eps.injected + i);
      e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w;
    } else {
      philox_normal4(eps, eps.row_offset + (uint64_t)b, outer, (uint32_t)(n >> 2), e);
    }
    const float4 g = *reinterpret_cast<const float4*>(dy + i), s = *reinterpret_cast<const float4*>(sd + i);
    *reinterpret_cast<float4*>(dV + i) =
        make_float4(g.x * e[0] / (2.f * s.x), g.y * e[1] / (2.f * s.y), g.z * e[2] / (2.f * s.z), g.w * e[3] / (2.f * s.w));
    if (dm) {
      float4 zz = make_float4(1.f, 1.f, 1.f, 1.f);
      if (z) zz = *reinterpret_cast<const float4*>(z + b * N + n);
      *reinterpret_cast<float4*>(dm + i) = make_float4(g.x * zz.x, g.y * zz.y, g.z * zz.z, g.w * zz.w);
    }
  }
}
static int dy_split(mnf_ctx* ctx, int64_t M, int N, int64_t HoWo, const float* dy, const float* sd, const float* z,
                    const NoiseDev& eps, float* dV, float* dm) {
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (N % 4 == 0 && al(dy) && al(sd) && al(z) && al(dV) && al(dm) && al(eps.injected))
    dy_split_kernel4<<<ew_grid(ctx, M * (N / 4)), 256, 0, ctx->stream>>>(M, N, HoWo, dy, sd, z, eps, dV, dm);
  else
    dy_split_kernel<<<ew_grid(ctx, M * N), 256, 0, ctx->stream>>>(M, N, HoWo, dy, sd, z, eps, dV, dm);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

// out0[n] = sum_m a0[m,n], out1[n] = sum_m a1[m,n]   (block = 32 columns x 8 row lanes).
// The rows are split over gridDim.y CTAs (a conv layer's bias gradient sums B*Ho*Wo rows of a
// 20- or 50-column matrix: one CTA per 32 columns walked 73,728 rows alone, 3.4 ms); each writes
// its partial sums to part[split][2][N] and colsum2_fold adds them in a fixed order.
#define COLSUM_SPLITS 256
__global__ void colsum2_kernel(int64_t M, int N, const float* __restrict__ a0, const float* __restrict__ a1,
                               float* __restrict__ part) {
  __shared__ float s0[8][33], s1[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + cx;
  const int64_t chunk = (M + gridDim.y - 1) / gridDim.y;
  const int64_t mlo = blockIdx.y * chunk, mhi = (mlo + chunk < M) ? mlo + chunk : M;
  float t0 = 0.f, t1 = 0.f;
  if (n < N)
    for (int64_t m = mlo + ry; m < mhi; m += 8) {
      t0 += a0[m * N + n];
      if (a1) t1 += a1[m * N + n];
    }
  s0[ry][cx] = t0;
  s1[ry][cx] = t1;
  __syncthreads();
  if (ry == 0 && n < N) {
    for (int r = 1; r < 8; ++r) { t0 += s0[r][cx]; t1 += s1[r][cx]; }
    part[((int64_t)blockIdx.y * 2 + 0) * N + n] = t0;
    part[((int64_t)blockIdx.y * 2 + 1) * N + n] = t1;
  }
}
__global__ void __launch_bounds__(256) colsum2_fold(int N, int splits, const float* __restrict__ part,
                                                    float* __restrict__ out0, float* __restrict__ out1) {
  // a warp per column: lanes stride over the row splits, then a butterfly (fixed order, and 8 dependent
  // loads per lane instead of a 256-long serial chain in one thread)
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  float t0 = 0.f, t1 = 0.f;
  for (int sp = lane; sp < splits; sp += 32) {
    t0 += part[((int64_t)sp * 2 + 0) * N + n];
    t1 += part[((int64_t)sp * 2 + 1) * N + n];
  }
  t0 = warp_sum(t0);
  t1 = warp_sum(t1);
  if (lane == 0) {
    if (out0) out0[n] = t0;
    if (out1) out1[n] = t1;
  }
}
static inline size_t colsum2_scratch_floats(int64_t N) { return (size_t)COLSUM_SPLITS * 2 * (size_t)N; }
// part: colsum2_scratch_floats(N) floats of scratch
static int colsum2(mnf_ctx* ctx, int64_t M, int N, const float* a0, const float* a1, float* out0, float* out1,
                   float* part) {
  int splits = (int)cdiv(M, 64);
  if (splits > COLSUM_SPLITS) splits = COLSUM_SPLITS;
  if (splits < 1) splits = 1;
  colsum2_kernel<<<dim3((unsigned)cdiv(N, 32), (unsigned)splits), 256, 0, ctx->stream>>>(M, N, a0, a1, part);
  MNF_LAUNCHED(ctx);
  colsum2_fold<<<(unsigned)cdiv(N, 8), 256, 0, ctx->stream>>>(N, splits, part, out0, a1 ? out1 : nullptr);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

// d_log_std_W = in_range ? 2 Vw dVw : 0 (in place on g); d_log_var_b = in_range ? vb dvb : 0
__global__ void clip_grad_kernel(int64_t nW, const float* __restrict__ Ls, const float* __restrict__ Vw, float max_std,
                                 float* __restrict__ gW, int N, const float* __restrict__ lvb,
                                 const float* __restrict__ vb, const float* __restrict__ dvb, float* __restrict__ gb) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nW; i += (int64_t)gridDim.x * blockDim.x) {
    if (gW) gW[i] = expf(Ls[i]) <= max_std ? 2.f * Vw[i] * gW[i] : 0.f;
    if (gb && i < N) gb[i] = expf(lvb[i]) <= max_std * max_std ? vb[i] * dvb[i] : 0.f;
  }
}

// ------------------------------------------------------------------------- dense backward
__global__ void dense_ax_kernel(int64_t n, const float* __restrict__ x, const float* __restrict__ z,
                                float* __restrict__ A, float* __restrict__ X2) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i];
    A[i] = z ? v * z[i] : v;
    X2[i] = v * v;
  }
}
__global__ void dense_dx_kernel(int64_t n, const float* __restrict__ x, const float* __restrict__ z,
                                const float* __restrict__ dA, const float* __restrict__ dX2, float* __restrict__ dx,
                                float* __restrict__ dz) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = x[i], a = dA[i];
    if (dz) dz[i] = a * v;
    if (dx) dx[i] = a * (z ? z[i] : 1.f) + 2.f * v * dX2[i];
  }
}

extern "C" int mnf_dense_backward(mnf_ctx* ctx, int64_t B, int64_t K, int64_t N, const float* x, const float* z,
                                  const float* mean_W, const float* log_std_W, const float* log_var_b,
                                  float max_std, const mnf_noise* eps, const float* sd, const float* dy, float* dx,
                                  float* dz, float* d_mean_W, float* d_log_std_W, float* d_mean_b,
                                  float* d_log_var_b) {
  MNF_CHECK_ARG(ctx && x && mean_W && log_std_W && log_var_b && eps && sd && dy, "mnf_dense_backward: NULL argument");
  MNF_CHECK_ARG(B >= 1 && K >= 1 && N >= 1 && K < (1 << 30) && N < (1 << 30) && B < (1ll << 31),
                "mnf_dense_backward: bad shape");
  const int64_t KN = K * N, BK = B * K, BN = B * N;
  size_t bytes = pad256(4 * KN) + pad256(4 * N) * 2 + pad256(4 * BN) + 4 * pad256(4 * BK) + 4096 +
                 pad256(4 * colsum2_scratch_floats(N));
  char* ws = (char*)mnf_workspace(ctx, bytes);
  if (!ws) return MNF_ERR_CUDA;
  Bump bp = {ws, 0};
  float* Vw = bp.take<float>(KN);
  float* vb = bp.take<float>(N);
  float* dvb = bp.take<float>(N);
  float* dV = bp.take<float>(BN);
  float* A = bp.take<float>(BK);
  float* X2 = bp.take<float>(BK);
  float* dA = bp.take<float>(BK);
  float* dX2 = bp.take<float>(BK);
  float* csum = bp.take<float>(colsum2_scratch_floats(N));
  bwd_prep_var<<<ew_grid(ctx, KN), 256, 0, ctx->stream>>>(KN, log_std_W, max_std, Vw, (int)N, log_var_b, vb);
  MNF_LAUNCHED(ctx);
  {
    int rc0 = dy_split(ctx, B, (int)N, 1, dy, sd, nullptr, make_noise(eps), dV, nullptr);
    if (rc0) return rc0;
  }
  int rc;
  if (d_mean_b || d_log_var_b) {
    if ((rc = colsum2(ctx, B, (int)N, dy, dV, d_mean_b, dvb, csum))) return rc;
  }
  if (dx || dz) {
    // dA = dy @ mean_W^T ; dX2 = dV @ Vw^T
    if ((rc = gemm(ctx, (int)B, (int)K, (int)N, dy, N, 1, mean_W, 1, N, dA, K, false))) return rc;
    if ((rc = gemm(ctx, (int)B, (int)K, (int)N, dV, N, 1, Vw, 1, N, dX2, K, false))) return rc;
    dense_dx_kernel<<<ew_grid(ctx, BK), 256, 0, ctx->stream>>>(BK, x, z, dA, dX2, dx, dz);
    MNF_LAUNCHED(ctx);
  }
  if (d_mean_W || d_log_std_W) {
    dense_ax_kernel<<<ew_grid(ctx, BK), 256, 0, ctx->stream>>>(BK, x, z, A, X2);
    MNF_LAUNCHED(ctx);
    if (d_mean_W && (rc = gemm(ctx, (int)K, (int)N, (int)B, A, 1, K, dy, N, 1, d_mean_W, N, false))) return rc;
    if (d_log_std_W && (rc = gemm(ctx, (int)K, (int)N, (int)B, X2, 1, K, dV, N, 1, d_log_std_W, N, false))) return rc;
  }
  if (d_log_std_W || d_log_var_b) {
    clip_grad_kernel<<<ew_grid(ctx, KN), 256, 0, ctx->stream>>>(KN, log_std_W, Vw, max_std, d_log_std_W, (int)N,
                                                               log_var_b, vb, dvb, d_log_var_b);
    MNF_LAUNCHED(ctx);
  }
  return MNF_OK;
}

// -------------------------------------------------------------------------- conv backward
struct ConvGeo {
  int64_t B;
  int H, W, C, kh, kw, F, Ho, Wo, stride, pad_t, pad_l;
};

// cols[m, (i,j,c)] = x[b, ho*s+i-pt, wo*s+j-pl, c] (0 outside); m local to the chunk.  The variance twin squares it inside the GEMM.
__global__ void im2col_kernel(ConvGeo g, int64_t b0, int64_t Mc, const float* __restrict__ x, float* __restrict__ cols) {
  // a warp per output pixel: the 64-bit index arithmetic is done once per row, the lanes walk the
  // patch (a filter row is kw*C contiguous input elements) with 32-bit divisions only
  const int Kc = g.kh * g.kw * g.C, kwC = g.kw * g.C, HoWo = g.Ho * g.Wo;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t m = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); m < Mc; m += (int64_t)gridDim.x * wpb) {
    const int64_t bl = m / HoWo;
    const int pix = (int)(m - bl * HoWo), ho = pix / g.Wo, wo = pix - ho * g.Wo;
    const int hi0 = ho * g.stride - g.pad_t, wi0 = wo * g.stride - g.pad_l;
    const float* xb = x + (b0 + bl) * (int64_t)g.H * g.W * g.C;
    float* dst = cols + m * Kc;
    for (int k = lane; k < Kc; k += 32) {
      const int i = k / kwC, rem = k - i * kwC, j = rem / g.C;
      const int hi = hi0 + i, wi = wi0 + j;
      float v = 0.f;
      if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.W) v = __ldg(xb + ((int64_t)hi * g.W + wi0) * g.C + rem);
      dst[k] = v;
    }
  }
}

// dx[b,hi,wi,c] = sum_{i,j} dcols[m(b,ho,wo),(i,j,c)] + 2 x * sum dcols2[...], gather form (no atomics)
// IDX = int when the chunk's element count fits (64-bit divisions cost ~100 instructions each)
template <typename IDX>
__global__ void col2im_kernel(ConvGeo g, int64_t b0, int64_t Bc, const float* __restrict__ x,
                              const float* __restrict__ dcols, const float* __restrict__ dcols2,
                              float* __restrict__ dx) {
  const int Kc = g.kh * g.kw * g.C;
  const IDX total = (IDX)(Bc * g.H * g.W * g.C);
  for (IDX idx = (IDX)blockIdx.x * (IDX)blockDim.x + (IDX)threadIdx.x; idx < total; idx += (IDX)gridDim.x * (IDX)blockDim.x) {
    int c = (int)(idx % g.C);
    IDX t = idx / g.C;
    int wi = (int)(t % g.W);
    t /= g.W;
    int hi = (int)(t % g.H);
    int64_t bl = (int64_t)(t / g.H);
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < g.kh; ++i) {
      int hh = hi + g.pad_t - i;
      if (hh < 0 || hh % g.stride) continue;
      int ho = hh / g.stride;
      if (ho >= g.Ho) continue;
      for (int j = 0; j < g.kw; ++j) {
        int ww = wi + g.pad_l - j;
        if (ww < 0 || ww % g.stride) continue;
        int wo = ww / g.stride;
        if (wo >= g.Wo) continue;
        int64_t m = (bl * g.Ho + ho) * g.Wo + wo;
        int64_t o = m * Kc + (i * g.kw + j) * g.C + c;
        s1 += dcols[o];
        s2 += dcols2[o];
      }
    }
    int64_t gi = (b0 * g.H * g.W) * g.C + idx;
    dx[gi] = s1 + 2.f * x[gi] * s2;
  }
}

// dz[b,f] = sum_pix dy * mean   (one block per image, threads over f)
__global__ void __launch_bounds__(256) conv_dz_kernel(int64_t B, int64_t HoWo, int F, const float* __restrict__ dy,
                                                      const float* __restrict__ mean, float* __restrict__ dz) {
  // one CTA per image.  F <= 256: the threads form G = 256 / F pixel groups x F filters (all lanes
  // busy, consecutive threads on consecutive addresses), partial sums folded through shared memory;
  // wider layers keep a thread per filter.
  __shared__ float red[256];
  const int64_t b = blockIdx.x;
  const float* dyb = dy + b * HoWo * F;
  const float* mb = mean + b * HoWo * F;
  if (F <= 256) {
    const int G = 256 / F, f = threadIdx.x % F, g = threadIdx.x / F;
    float s = 0.f;
    if (g < G)
      for (int64_t p = g; p < HoWo; p += G) s = fmaf(dyb[p * F + f], mb[p * F + f], s);
    red[threadIdx.x] = g < G ? s : 0.f;
    __syncthreads();
    if (threadIdx.x < F) {
      for (int gg = 1; gg < G; ++gg) s += red[gg * F + threadIdx.x];
      dz[b * F + threadIdx.x] = s;
    }
    return;
  }
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float s = 0.f;
    for (int64_t p = 0; p < HoWo; ++p) s = fmaf(dyb[p * F + f], mb[p * F + f], s);
    dz[b * F + f] = s;
  }
}

extern "C" int mnf_conv2d_backward(mnf_ctx* ctx, const mnf_conv_shape* s, const float* x, const float* z,
                                   const float* mean_W, const float* log_std_W, const float* log_var_b,
                                   float max_std, const mnf_noise* eps, const float* mean_saved, const float* sd,
                                   const float* dy, float* dx, float* dz, float* d_mean_W, float* d_log_std_W,
                                   float* d_mean_b, float* d_log_var_b) {
  MNF_CHECK_ARG(ctx && s && x && mean_W && log_std_W && log_var_b && eps && sd && dy, "mnf_conv2d_backward: NULL argument");
  MNF_CHECK_ARG(!dz || mean_saved, "mnf_conv2d_backward: dz needs mean_saved");
  int64_t Ho, Wo;
  int rc = mnf_conv_out_hw(s, &Ho, &Wo);
  if (rc) return rc;
  MNF_CHECK_ARG(s->B >= 1, "mnf_conv2d_backward: empty batch");
  ConvGeo g;
  g.B = s->B; g.H = (int)s->H; g.W = (int)s->W; g.C = (int)s->C; g.kh = (int)s->kh; g.kw = (int)s->kw;
  g.F = (int)s->F; g.Ho = (int)Ho; g.Wo = (int)Wo; g.stride = (int)s->stride; g.pad_t = g.pad_l = 0;
  if (s->same) {
    int64_t ph = (Ho - 1) * s->stride + s->kh - s->H, pw = (Wo - 1) * s->stride + s->kw - s->W;
    g.pad_t = (int)((ph > 0 ? ph : 0) / 2);
    g.pad_l = (int)((pw > 0 ? pw : 0) / 2);
  }
  const int F = g.F;
  const int64_t Kc = (int64_t)g.kh * g.kw * g.C, HoWo = Ho * Wo, M = s->B * HoWo, KF = Kc * F, MF = M * F;
  // images per chunk: four [Mc,Kc] buffers within ~1 GiB
  int64_t Bc = (int64_t)(((size_t)1 << 30) / (4 * 4 * (size_t)(HoWo * Kc)));
  if (Bc < 1) Bc = 1;
  if (Bc > s->B) Bc = s->B;
  const int64_t Mc_max = Bc * HoWo;
  MNF_CHECK_ARG(Mc_max < (1ll << 31) && Kc < (1 << 30), "mnf_conv2d_backward: shape too large");
  size_t bytes = pad256(4 * KF) * 2 + pad256(4 * F) * 2 + pad256(4 * MF) * 2 + 3 * pad256(4 * Mc_max * Kc) + 4096 +
                 pad256(4 * colsum2_scratch_floats(F));
  char* ws = (char*)mnf_workspace(ctx, bytes);
  if (!ws) return MNF_ERR_CUDA;
  Bump bp = {ws, 0};
  float* Vw = bp.take<float>(KF);
  float* dVw = bp.take<float>(KF);
  float* vb = bp.take<float>(F);
  float* dvb = bp.take<float>(F);
  float* dm = bp.take<float>(MF);
  float* dv = bp.take<float>(MF);
  float* cols = bp.take<float>(Mc_max * Kc);
  float* dcols = bp.take<float>(Mc_max * Kc);
  float* dcols2 = bp.take<float>(Mc_max * Kc);
  float* csum = bp.take<float>(colsum2_scratch_floats(F));
  bwd_prep_var<<<ew_grid(ctx, KF), 256, 0, ctx->stream>>>(KF, log_std_W, max_std, Vw, F, log_var_b, vb);
  MNF_LAUNCHED(ctx);
  if ((rc = dy_split(ctx, M, F, HoWo, dy, sd, z, make_noise(eps), dv, dm))) return rc;
  if (d_mean_b || d_log_var_b) {
    if ((rc = colsum2(ctx, M, F, dm, dv, d_mean_b, dvb, csum))) return rc;
  }
  if (dz) {
    conv_dz_kernel<<<(unsigned)s->B, 256, 0, ctx->stream>>>(s->B, HoWo, F, dy, mean_saved, dz);
    MNF_LAUNCHED(ctx);
  }
  const bool wgrad = d_mean_W || d_log_std_W;
  if (d_mean_W) MNF_CUDA(cudaMemsetAsync(d_mean_W, 0, sizeof(float) * KF, ctx->stream));
  if (d_log_std_W) MNF_CUDA(cudaMemsetAsync(dVw, 0, sizeof(float) * KF, ctx->stream));
  for (int64_t b0 = 0; b0 < s->B && (wgrad || dx); b0 += Bc) {
    const int64_t bc = (s->B - b0 < Bc) ? s->B - b0 : Bc;
    const int64_t Mc = bc * HoWo;
    const float* dmc = dm + b0 * HoWo * F;
    const float* dvc = dv + b0 * HoWo * F;
    if (wgrad) {
      im2col_kernel<<<ew_grid(ctx, Mc * Kc), 256, 0, ctx->stream>>>(g, b0, Mc, x, cols);
      MNF_LAUNCHED(ctx);
      // dW += cols^T @ dm   (M=Kc, N=F, K=Mc)
      if (d_mean_W && (rc = gemm(ctx, (int)Kc, F, (int)Mc, cols, 1, Kc, dmc, F, 1, d_mean_W, F, true))) return rc;
      // the variance twin squares the same patches on the fly (no second im2col matrix)
      if (d_log_std_W && (rc = gemm(ctx, (int)Kc, F, (int)Mc, cols, 1, Kc, dvc, F, 1, dVw, F, true, true))) return rc;
    }
    if (dx) {
      // dcols = dm @ W^T  (M=Mc, N=Kc, K=F): B(k=f, n=kc) = W[kc*F + f]
      if ((rc = gemm(ctx, (int)Mc, (int)Kc, F, dmc, F, 1, mean_W, 1, F, dcols, Kc, false))) return rc;
      if ((rc = gemm(ctx, (int)Mc, (int)Kc, F, dvc, F, 1, Vw, 1, F, dcols2, Kc, false))) return rc;
      if (bc * g.H * g.W * g.C < (1ll << 30))
        col2im_kernel<int><<<ew_grid(ctx, bc * g.H * g.W * g.C), 256, 0, ctx->stream>>>(g, b0, bc, x, dcols, dcols2, dx);
      else
        col2im_kernel<int64_t><<<ew_grid(ctx, bc * g.H * g.W * g.C), 256, 0, ctx->stream>>>(g, b0, bc, x, dcols, dcols2, dx);
      MNF_LAUNCHED(ctx);
    }
  }
  if (d_log_std_W) MNF_CUDA(cudaMemcpyAsync(d_log_std_W, dVw, sizeof(float) * KF, cudaMemcpyDeviceToDevice, ctx->stream));
  if (d_log_std_W || d_log_var_b) {
    clip_grad_kernel<<<ew_grid(ctx, KF), 256, 0, ctx->stream>>>(KF, log_std_W, Vw, max_std, d_log_std_W, F, log_var_b,
                                                               vb, dvb, d_log_var_b);
    MNF_LAUNCHED(ctx);
  }
  return MNF_OK;
}

// --------------------------------------------------------------------------- flow backward
#define FB_ROWS 32
#define FB_THREADS 256
#define FB_HMAX 64
#define FB_JMAX (FB_HMAX / 8)

struct FlowBwdArgs {
  int64_t B;
  int D, H, parity;
  const float* xin;   // states[k]
  const float* hid;   // saved hidden activations [B,H]
  const float* dy;    // grad wrt states[k+1]
  const float* dld;   // [B] or nullptr
  const float* add;   // d_states[k] or nullptr (added to the output)
  float* dx;          // grad wrt states[k]
  const float *w1, *ws, *bs, *wt, *bt, *m1, *m2;
  int64_t ld2;
  NoiseDev bern;
  float *gw1, *gb1, *gws, *gbs, *gwt, *gbt;
  float* dhbuf;  // [B,H] zeroed scratch: dL/dh summed over the column splits (PHASE 1 -> PHASE 2)
  int chunks_per_split;
};

// PHASE 0: the whole step in one CTA per 32 rows.  With few rows (a 128-row minibatch, or the single
// row of kl_div's chains) that is 1-4 CTAs walking D/32 barrier-separated column chunks twice
// (0.4-0.6 ms at D = 800): PHASE 1 / PHASE 2 run stage A resp. stages B + C with the column chunks
// spread over blockIdx.y, exchanging dL/dh through dhbuf (atomic adds, like the weight gradients).
template <int KIND, int PHASE>
__global__ void __launch_bounds__(FB_THREADS) flow_mlp_step_bwd(FlowBwdArgs a) {
  __shared__ float hs[FB_ROWS][FB_HMAX + 1];
  __shared__ float dps[FB_ROWS][FB_HMAX + 1];
  __shared__ float dss[FB_ROWS][33], dts[FB_ROWS][33];
  __shared__ float wsm[2][FB_HMAX][33];
  const int tid = threadIdx.x, r = tid >> 3, g = tid & 7;
  const int64_t row0 = (int64_t)blockIdx.x * FB_ROWS, row = row0 + r;
  const bool row_ok = row < a.B;
  const int D = a.D, H = a.H;
  const int c_lo = PHASE == 0 ? 0 : (int)blockIdx.y * a.chunks_per_split * 32;
  const int c_hi = PHASE == 0 ? D : ((c_lo + a.chunks_per_split * 32 < D) ? c_lo + a.chunks_per_split * 32 : D);
  float (*xs)[33] = dss;  // stage C reuses dss for the input chunk

  for (int e = tid; e < FB_ROWS * H; e += FB_THREADS) {
    int rr = e / H, h = e - rr * H;
    hs[rr][h] = (row0 + rr < a.B) ? a.hid[(row0 + rr) * H + h] : 0.f;
  }
  const float dldv = (row_ok && a.dld) ? a.dld[row] : 0.f;
  float dh[FB_JMAX];
#pragma unroll
  for (int j = 0; j < FB_JMAX; ++j) dh[j] = 0.f;
  __syncthreads();

  // ---- stage A: recompute (s,t), form (ds,dt), direct dx term, dh, head weight grads
  for (int c0 = c_lo; c0 < c_hi && PHASE != 2; c0 += 32) {
    for (int e = tid; e < H * 32; e += FB_THREADS) {
      int h = e >> 5, cc = e & 31, c = c0 + cc;
      float vs = 0.f, vt = 0.f;
      if (c < D) {
        vs = a.ws[(int64_t)h * a.ld2 + c];
        vt = a.wt[(int64_t)h * a.ld2 + c];
        if (a.m2) { vs *= a.m2[(int64_t)h * a.ld2 + c]; vt *= a.m2[(int64_t)h * a.ld2 + D + c]; }
      }
      wsm[0][h][cc] = vs;
      wsm[1][h][cc] = vt;
    }
    __syncthreads();
    float s[4] = {0.f, 0.f, 0.f, 0.f}, t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int h = 0; h < H; ++h) {
      float hv = hs[r][h];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        s[q] = fmaf(hv, wsm[0][h][g + 8 * q], s[q]);
        t[q] = fmaf(hv, wsm[1][h][g + 8 * q], t[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int cc = g + 8 * q, c = c0 + cc;
      float ds = 0.f, dt = 0.f;
      if (row_ok && c < D) {
        float sv = s[q] + a.bs[c], tv = t[q] + a.bt[c];
        float xv = a.xin[row * D + c];
        float dxd;
        if (KIND == MNF_FLOW_IAF) {
          float dyv = a.dy[row * D + (a.parity ? D - 1 - c : c)];
          float es = expf(sv);
          ds = dyv * xv * es + dldv;
          dt = dyv;
          dxd = dyv * es;
        } else {
          float dyv = a.dy[row * D + c];
          float m = noise_bern1(a.bern, row, (uint32_t)c, D, 0.5f);
          float gate = 1.f / (1.f + expf(-sv));
          float dg = dyv * ((1.f - m) * xv - tv) + dldv * (1.f - m) / gate;
          ds = dg * gate * (1.f - gate);
          dt = dyv * (1.f - gate);
          dxd = dyv * ((1.f - m) * gate + m);
        }
        a.dx[row * D + c] = dxd;
      }
      dss[r][cc] = ds;
      dts[r][cc] = dt;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < FB_JMAX; ++j) {
      int h = g + 8 * j;
      if (h < H) {
        float acc = dh[j];
#pragma unroll 8
        for (int cc = 0; cc < 32; ++cc) acc = fmaf(dss[r][cc], wsm[0][h][cc], fmaf(dts[r][cc], wsm[1][h][cc], acc));
        dh[j] = acc;
      }
    }
    for (int e = tid; e < H * 32; e += FB_THREADS) {
      int h = e >> 5, cc = e & 31, c = c0 + cc;
      if (c >= D) continue;
      float gs = 0.f, gt = 0.f;
#pragma unroll 8
      for (int rr = 0; rr < FB_ROWS; ++rr) {
        float hv = hs[rr][h];
        gs = fmaf(hv, dss[rr][cc], gs);
        gt = fmaf(hv, dts[rr][cc], gt);
      }
      if (a.m2) { gs *= a.m2[(int64_t)h * a.ld2 + c]; gt *= a.m2[(int64_t)h * a.ld2 + D + c]; }
      atomicAdd(a.gws + (int64_t)h * a.ld2 + c, gs);
      atomicAdd(a.gwt + (int64_t)h * a.ld2 + c, gt);
    }
    if (tid < 64) {
      int cc = tid & 31, c = c0 + cc;
      if (c < D) {
        float acc = 0.f;
        for (int rr = 0; rr < FB_ROWS; ++rr) acc += (tid < 32) ? dss[rr][cc] : dts[rr][cc];
        atomicAdd((tid < 32 ? a.gbs : a.gbt) + c, acc);
      }
    }
    __syncthreads();
  }

  if (PHASE == 1) {
#pragma unroll
    for (int j = 0; j < FB_JMAX; ++j) {
      int h = g + 8 * j;
      if (h < H && row_ok) atomicAdd(a.dhbuf + row * H + h, dh[j]);
    }
    return;
  }
  if (PHASE == 2) {
#pragma unroll
    for (int j = 0; j < FB_JMAX; ++j) {
      int h = g + 8 * j;
      dh[j] = (h < H && row_ok) ? a.dhbuf[row * H + h] : 0.f;
    }
  }

  // ---- stage B: through the hidden activation
#pragma unroll
  for (int j = 0; j < FB_JMAX; ++j) {
    int h = g + 8 * j;
    if (h < H) {
      float hv = hs[r][h];
      float d = (KIND == MNF_FLOW_IAF) ? (hv > 0.f ? dh[j] : 0.f) : dh[j] * (1.f - hv * hv);
      dps[r][h] = row_ok ? d : 0.f;
    }
  }
  __syncthreads();
  if (tid < H && (PHASE == 0 || blockIdx.y == 0)) {
    float acc = 0.f;
    for (int rr = 0; rr < FB_ROWS; ++rr) acc += dps[rr][tid];
    atomicAdd(a.gb1 + tid, acc);
  }

  // ---- stage C: dW1 and the MLP-path term of dx
  for (int d0 = c_lo; d0 < c_hi; d0 += 32) {
    __syncthreads();
    for (int e = tid; e < 32 * H; e += FB_THREADS) {
      int dd = e / H, h = e - dd * H, d = d0 + dd;
      float w = 0.f;
      if (d < D) {
        w = a.w1[(int64_t)d * H + h];
        if (a.m1) w *= a.m1[(int64_t)d * H + h];
      }
      wsm[0][h][dd] = w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = tid + FB_THREADS * q, rr = e >> 5, dd = e & 31, d = d0 + dd;
      float v = 0.f;
      if (row0 + rr < a.B && d < D) {
        v = a.xin[(row0 + rr) * D + d];
        if (KIND == MNF_FLOW_RNVP) v *= noise_bern1(a.bern, row0 + rr, (uint32_t)d, D, 0.5f);
      }
      xs[rr][dd] = v;
    }
    __syncthreads();
    if (row_ok) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int dd = g + 8 * q, d = d0 + dd;
        if (d < D) {
          float acc = 0.f;
          for (int h = 0; h < H; ++h) acc = fmaf(dps[r][h], wsm[0][h][dd], acc);
          if (KIND == MNF_FLOW_RNVP) acc *= noise_bern1(a.bern, row, (uint32_t)d, D, 0.5f);
          float out = a.dx[row * D + d] + acc;
          if (a.add) out += a.add[row * D + d];
          a.dx[row * D + d] = out;
        }
      }
    }
    for (int e = tid; e < 32 * H; e += FB_THREADS) {
      int dd = e / H, h = e - dd * H, d = d0 + dd;
      if (d >= D) continue;
      float acc = 0.f;
#pragma unroll 8
      for (int rr = 0; rr < FB_ROWS; ++rr) acc = fmaf(xs[rr][dd], dps[rr][h], acc);
      if (a.m1) acc *= a.m1[(int64_t)d * H + h];
      atomicAdd(a.gw1 + (int64_t)d * H + h, acc);
    }
  }
}

// planar: per-row pass (32 rows per CTA), then a parameter-level finish
struct PlanarBwdArgs {
  int64_t B;
  int D;
  const float* zin;
  const float* dy;
  const float* dld;
  const float* add;
  float* dx;
  const float *w, *b, *u;
  const float* scal;  // k, w.u_hat
  float* tmp;         // [2D + 2]: du_hat[D], dw_direct[D], dwu, db
};

__global__ void __launch_bounds__(256) planar_step_bwd(PlanarBwdArgs a) {
  __shared__ float th_s[32], da_s[32];
  __shared__ float red[2][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * 32;
  const int D = a.D;
  const float k = a.scal[0], wu = a.scal[1];
  float dwu_p = 0.f, db_p = 0.f;
  for (int rr = wid; rr < 32; rr += 8) {
    int64_t row = row0 + rr;
    float th = 0.f, da = 0.f;
    if (row < a.B) {
      float dot = 0.f, dxu = 0.f;
      for (int d = lane; d < D; d += 32) {
        float wv = a.w[d];
        dot = fmaf(a.zin[row * D + d], wv, dot);
        dxu = fmaf(a.dy[row * D + d], a.u[d] + k * wv, dxu);
      }
      dot = warp_sum(dot);
      dxu = warp_sum(dxu);
      th = tanhf(dot + a.b[0]);
      float dth2 = 1.f - th * th;
      float dq = (a.dld ? a.dld[row] : 0.f) / (1.f + dth2 * wu);
      float dth = dxu + dq * wu * (-2.f * th);
      da = dth * dth2;
      if (lane == 0) { dwu_p += dq * dth2; db_p += da; }
    }
    if (lane == 0) { th_s[rr] = th; da_s[rr] = da; }
  }
  if (lane == 0) { red[0][wid] = dwu_p; red[1][wid] = db_p; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int i = 0; i < 8; ++i) { s0 += red[0][i]; s1 += red[1][i]; }
    atomicAdd(a.tmp + 2 * D, s0);
    atomicAdd(a.tmp + 2 * D + 1, s1);
  }
  const int nrows = (int)((a.B - row0 < 32) ? a.B - row0 : 32);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float duh = 0.f, dwd = 0.f, wv = a.w[d];
    for (int rr = 0; rr < nrows; ++rr) {
      int64_t o = (row0 + rr) * D + d;
      float dyv = a.dy[o];
      duh = fmaf(dyv, th_s[rr], duh);
      dwd = fmaf(a.zin[o], da_s[rr], dwd);
      float out = dyv + da_s[rr] * wv;
      if (a.add) out += a.add[o];
      a.dx[o] = out;
    }
    atomicAdd(a.tmp + d, duh);
    atomicAdd(a.tmp + D + d, dwd);
  }
}

__global__ void planar_finish_bwd(int D, const float* __restrict__ w, const float* __restrict__ u,
                                  const float* __restrict__ tmp, float* gw, float* gu, float* gb) {
  __shared__ double sh[3][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  auto bsum = [&](double v, int slot) {
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) sh[slot][wid] = v;
    __syncthreads();
    double t = 0;
    for (int i = 0; i < nw; ++i) t += sh[slot][i];
    return t;
  };
  double puw = 0, pww = 0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    puw += (double)u[d] * w[d];
    pww += (double)w[d] * w[d];
  }
  const float uw = (float)bsum(puw, 0), ww = (float)bsum(pww, 1);
  const float sp = uw > 20.f ? uw : log1pf(expf(uw));
  const float suw = -1.f + sp;
  const float k = (suw - uw) / ww;
  const float dwu = tmp[2 * D];
  double pdk = 0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) pdk += (double)(tmp[d] + dwu * w[d]) * w[d];
  const float dk = (float)bsum(pdk, 2);
  const float dww = -dk * (suw - uw) / (ww * ww);
  const float sig = 1.f / (1.f + expf(-uw));
  const float duw = dk / ww * (sig - 1.f);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float duh = tmp[d] + dwu * w[d];
    float uhat = u[d] + k * w[d];
    gu[d] += duh + duw * w[d];
    gw[d] += dwu * uhat + tmp[D + d] + k * duh + duw * u[d] + dww * 2.f * w[d];
  }
  if (threadIdx.x == 0) gb[0] += tmp[2 * D + 1];
}

__global__ void planar_prep_bwd(int D, const float* __restrict__ w, const float* __restrict__ u, float* scal);

extern "C" int mnf_flow_backward(mnf_ctx* ctx, int64_t B, int64_t D, int32_t T, const mnf_flow_step* steps,
                                 const float* states, const float* saved, const float* d_states, const float* d_zT,
                                 const float* d_log_dets, float* d_z0, const mnf_flow_step* gsteps) {
  MNF_CHECK_ARG(ctx && states && d_z0, "mnf_flow_backward: NULL argument");
  MNF_CHECK_ARG(B >= 1 && D >= 1 && D < (1 << 30) && T >= 0 && T <= 64, "mnf_flow_backward: bad shape");
  MNF_CHECK_ARG(T == 0 || (steps && gsteps), "mnf_flow_backward: steps/grad steps NULL");
  const int64_t BD = B * D;
  // running gradient buffers (ping-pong) + planar temporaries live in the workspace
  size_t bytes = 2 * pad256(4 * BD) + pad256(4 * (2 * D + 2)) * (T + 1) + pad256(8 * (T + 1)) + 4096 +
                 pad256(4 * (size_t)B * FB_HMAX);
  char* ws = (char*)mnf_workspace(ctx, bytes);
  if (!ws) return MNF_ERR_CUDA;
  Bump bp = {ws, 0};
  float* buf[2] = {bp.take<float>(BD), bp.take<float>(BD)};
  float* dhbuf = bp.take<float>((size_t)B * FB_HMAX);
  float* scal = bp.take<float>(2 * (T + 1));
  // g = d_states[T] + d_zT
  float* cur = buf[0];
  MNF_CUDA(cudaMemsetAsync(cur, 0, sizeof(float) * BD, ctx->stream));
  if (d_states) MNF_CUDA(cudaMemcpyAsync(cur, d_states + (int64_t)T * BD, sizeof(float) * BD, cudaMemcpyDeviceToDevice, ctx->stream));
  if (d_zT) {
    int rc = mnf_axpy(ctx, BD, 1.0f, d_zT, cur);
    if (rc) return rc;
  }
  // offsets of the packed `saved` buffer
  int64_t soff[65];
  soff[0] = 0;
  for (int k = 0; k < T; ++k)
    soff[k + 1] = soff[k] + B * mnf_flow_step_hidden_total(steps[k]);
  int which = 0;
  for (int k = T - 1; k >= 0; --k) {
    const mnf_flow_step& st = steps[k];
    const mnf_flow_step& gs = gsteps[k];
    float* out = (k == 0) ? d_z0 : buf[which ^ 1];
    const float* add = d_states ? d_states + (int64_t)k * BD : nullptr;
    if (mnf_flow_step_is_generic(st)) {
      int rc = mnf_flow_generic_backward(ctx, B, D, st, gs, states + (int64_t)k * BD, saved ? saved + soff[k] : nullptr, cur,
                                         d_log_dets, add, out);
      if (rc) return rc;
    } else if (st.kind == MNF_FLOW_IAF || st.kind == MNF_FLOW_RNVP) {
      MNF_CHECK_ARG(saved, "mnf_flow_backward: saved hidden activations required");
      MNF_CHECK_ARG(st.hidden >= 1 && st.hidden <= FB_HMAX, "flow step %d: hidden=%d not in [1,%d]", k, st.hidden, FB_HMAX);
      MNF_CHECK_ARG(gs.w1 && gs.b1 && gs.ws && gs.bs && gs.wt && gs.bt, "flow step %d: NULL grad pointer", k);
      FlowBwdArgs a;
      a.B = B; a.D = (int)D; a.H = st.hidden; a.parity = st.parity;
      a.xin = states + (int64_t)k * BD; a.hid = saved + soff[k]; a.dy = cur; a.dld = d_log_dets; a.add = add;
      a.dx = out;
      a.w1 = st.w1; a.ws = st.ws; a.bs = st.bs; a.wt = st.wt; a.bt = st.bt; a.m1 = st.m1; a.m2 = st.m2; a.ld2 = st.ld2;
      a.bern = make_noise(&st.bern);
      a.gw1 = (float*)gs.w1; a.gb1 = (float*)gs.b1; a.gws = (float*)gs.ws; a.gbs = (float*)gs.bs;
      a.gwt = (float*)gs.wt; a.gbt = (float*)gs.bt;
      const int grid = (int)cdiv(B, FB_ROWS), nchunks = (int)cdiv(D, 32);
      int nsplit = (2 * ctx->sm_count) / grid;  // column splits that fill the machine about twice
      if (nsplit > nchunks) nsplit = nchunks;
      if (nsplit >= 2) {
        a.chunks_per_split = (int)cdiv(nchunks, nsplit);
        nsplit = (int)cdiv(nchunks, a.chunks_per_split);
        a.dhbuf = dhbuf;
        MNF_CUDA(cudaMemsetAsync(dhbuf, 0, sizeof(float) * B * st.hidden, ctx->stream));
        const dim3 g2((unsigned)grid, (unsigned)nsplit);
        if (st.kind == MNF_FLOW_IAF) {
          flow_mlp_step_bwd<MNF_FLOW_IAF, 1><<<g2, FB_THREADS, 0, ctx->stream>>>(a);
          flow_mlp_step_bwd<MNF_FLOW_IAF, 2><<<g2, FB_THREADS, 0, ctx->stream>>>(a);
        } else {
          flow_mlp_step_bwd<MNF_FLOW_RNVP, 1><<<g2, FB_THREADS, 0, ctx->stream>>>(a);
          flow_mlp_step_bwd<MNF_FLOW_RNVP, 2><<<g2, FB_THREADS, 0, ctx->stream>>>(a);
        }
        ctx->launches++;
      } else {
        a.chunks_per_split = nchunks;
        a.dhbuf = nullptr;
        if (st.kind == MNF_FLOW_IAF) flow_mlp_step_bwd<MNF_FLOW_IAF, 0><<<grid, FB_THREADS, 0, ctx->stream>>>(a);
        else flow_mlp_step_bwd<MNF_FLOW_RNVP, 0><<<grid, FB_THREADS, 0, ctx->stream>>>(a);
      }
      MNF_LAUNCHED(ctx);
    } else if (st.kind == MNF_FLOW_PLANAR) {
      MNF_CHECK_ARG(gs.w1 && gs.b1 && gs.wt, "planar step %d: NULL grad pointer", k);
      float* tmp = bp.take<float>(2 * D + 2);
      MNF_CUDA(cudaMemsetAsync(tmp, 0, sizeof(float) * (2 * D + 2), ctx->stream));
      planar_prep_bwd<<<1, 256, 0, ctx->stream>>>((int)D, st.w1, st.wt, scal + 2 * k);
      MNF_LAUNCHED(ctx);
      PlanarBwdArgs a;
      a.B = B; a.D = (int)D; a.zin = states + (int64_t)k * BD; a.dy = cur; a.dld = d_log_dets; a.add = add; a.dx = out;
      a.w = st.w1; a.b = st.b1; a.u = st.wt; a.scal = scal + 2 * k; a.tmp = tmp;
      planar_step_bwd<<<(unsigned)cdiv(B, 32), 256, 0, ctx->stream>>>(a);
      MNF_LAUNCHED(ctx);
      planar_finish_bwd<<<1, 256, 0, ctx->stream>>>((int)D, st.w1, st.wt, tmp, (float*)gs.w1, (float*)gs.wt, (float*)gs.b1);
      MNF_LAUNCHED(ctx);
    } else {
      MNF_CHECK_ARG(false, "flow step %d: unknown kind %d", k, st.kind);
    }
    cur = out;
    which ^= 1;
  }
  if (T == 0) {
    MNF_CUDA(cudaMemcpyAsync(d_z0, cur, sizeof(float) * BD, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return MNF_OK;
}

// same scalars as the forward's planar_prep (k, w.u_hat)
__global__ void planar_prep_bwd(int D, const float* __restrict__ w, const float* __restrict__ u, float* scal) {
  __shared__ double red[3][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double uw = 0, ww = 0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    uw += (double)u[d] * w[d];
    ww += (double)w[d] * w[d];
  }
  uw = warp_sum_d(uw);
  ww = warp_sum_d(ww);
  if (lane == 0) { red[0][wid] = uw; red[1][wid] = ww; }
  __syncthreads();
  uw = ww = 0;
  for (int i = 0; i < nw; ++i) { uw += red[0][i]; ww += red[1][i]; }
  float fuw = (float)uw, fww = (float)ww;
  float sp = fuw > 20.f ? fuw : log1pf(expf(fuw));
  float k = ((-1.f + sp) - fuw) / fww;
  double wu = 0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) wu += (double)w[d] * (u[d] + k * w[d]);
  wu = warp_sum_d(wu);
  if (lane == 0) red[2][wid] = wu;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < nw; ++i) t += red[2][i];
    scal[0] = k;
    scal[1] = (float)t;
  }
}

// d q0_mean += sum_b dz0 ; d q0_log_var += sum_b dz0 * eps * 0.5 exp(lv/2)   (App. A-2)
__global__ void sample_z_bwd_kernel(int64_t B, int D, const float* __restrict__ dz0, const float* __restrict__ lv,
                                    NoiseDev eps, float* gm, float* glv) {
  __shared__ float s0[8][33], s1[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + cx;
  float t0 = 0.f, t1 = 0.f;
  // rows are split over gridDim.y (a handful of CTAs walking the whole batch was 45 us per call);
  // the outputs accumulate, so the slices combine with atomic adds
  const int64_t chunk = (B + gridDim.y - 1) / gridDim.y;
  const int64_t blo = blockIdx.y * chunk, bhi = (blo + chunk < B) ? blo + chunk : B;
  if (d < D)
    for (int64_t b = blo + ry; b < bhi; b += 8) {
      float g = dz0[b * D + d];
      t0 += g;
      t1 = fmaf(g, noise_normal1(eps, b, 0u, (uint32_t)d, 1, D), t1);
    }
  s0[ry][cx] = t0;
  s1[ry][cx] = t1;
  __syncthreads();
  if (ry == 0 && d < D) {
    for (int r = 1; r < 8; ++r) { t0 += s0[r][cx]; t1 += s1[r][cx]; }
    if (gm) atomicAdd(gm + d, t0);
    if (glv) atomicAdd(glv + d, t1 * 0.5f * expf(0.5f * lv[d]));
  }
}

extern "C" int mnf_sample_z_backward(mnf_ctx* ctx, int64_t B, int64_t D, const float* d_z0, const float* q0_log_var,
                                     const mnf_noise* eps_z, float* d_q0_mean, float* d_q0_log_var) {
  MNF_CHECK_ARG(ctx && d_z0 && q0_log_var && eps_z, "mnf_sample_z_backward: NULL argument");
  MNF_CHECK_ARG(B >= 1 && D >= 1 && D < (1 << 30), "mnf_sample_z_backward: bad shape");
  int64_t splits = cdiv(B, 32);
  const int64_t want = cdiv(2 * ctx->sm_count, cdiv(D, 32));
  if (splits > want) splits = want;
  if (splits < 1) splits = 1;
  sample_z_bwd_kernel<<<dim3((unsigned)cdiv(D, 32), (unsigned)splits), 256, 0, ctx->stream>>>(
      B, (int)D, d_z0, q0_log_var, make_noise(eps_z), d_q0_mean, d_q0_log_var);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}
