// Normalizing-flow steps as fused row-tile kernels (forward).
//   IAF   : maf.py:47-53 (MAF.inverse bound as IAF.forward, maf.py:63), MADE made.py:77-91,
//           MaskedDense made.py:28-30
//   RNVP  : rnvp.py:33-47
//   planar: planar.py:22-36
//   chain : flows/__init__.py:22-29;  z0 sampling: mnf_dense.py:93-96 / mnf_conv.py:105-108
// One launch per flow step.  A CTA owns FT_ROWS rows; the D->H->2D coupling MLP runs out of
// shared memory (weights streamed in 32-wide chunks), the per-row log-det is reduced with
// warp shuffles and never touches global memory until the final [B] store.
#include "common.cuh"

#define FT_ROWS 32
#define FT_THREADS 256
#define FT_DC 32    // reduction / column chunk
#define FT_HMAX 64  // hidden units supported by the fused kernel (static smem budget)
#define FT_JMAX (FT_HMAX / 8)

struct FlowStepArgs {
  int64_t B;
  int D, H;
  int parity;
  const float* zin;    // [B,D] or nullptr when sampling z0 in the prologue
  const float* q0_mean;
  const float* q0_log_var;
  NoiseDev eps;        // z0 noise
  float* z0_out;       // where the sampled z0 is stored (states[0]); may be nullptr
  float* zout;         // [B,D]
  const float* ld_in;  // [B] or nullptr (treated as zeros)
  float* ld_out;       // [B] or nullptr
  float* hid_save;     // [B,H] or nullptr
  const float *w1, *b1, *ws, *bs, *wt, *bt, *m1, *m2;
  int64_t ld2;
  NoiseDev bern;
};

__device__ __forceinline__ float load_zin(const FlowStepArgs& a, int64_t row, int d) {
  if (a.zin) return a.zin[row * a.D + d];
  float e = noise_normal1(a.eps, row, 0u, (uint32_t)d, 1, a.D);
  return a.q0_mean[d] + expf(0.5f * a.q0_log_var[d]) * e;
}

template <int KIND>
__global__ void __launch_bounds__(FT_THREADS) flow_mlp_step_fwd(FlowStepArgs a) {
  __shared__ float zs[FT_ROWS][FT_DC + 1];
  __shared__ float wsm[2][FT_DC > FT_HMAX ? FT_DC : FT_HMAX][FT_DC + 1];  // weight chunks
  __shared__ float hs[FT_ROWS][FT_HMAX + 1];

  const int tid = threadIdx.x;
  const int r = tid >> 3, g = tid & 7;
  const int64_t row0 = (int64_t)blockIdx.x * FT_ROWS;
  const int64_t row = row0 + r;
  const bool row_ok = row < a.B;
  const int D = a.D, H = a.H;

  // ---- stage 1: h = act(in @ W1 + b1), in = z (IAF) or m*z (RNVP)
  float acc[FT_JMAX];
#pragma unroll
  for (int j = 0; j < FT_JMAX; ++j) acc[j] = 0.f;

  for (int d0 = 0; d0 < D; d0 += FT_DC) {
    // z chunk [32 rows][32 d]: thread loads 4 elements
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = tid + FT_THREADS * q;
      int rr = e >> 5, dd = e & 31;
      int64_t grow = row0 + rr;
      int d = d0 + dd;
      float v = 0.f;
      if (grow < a.B && d < D) {
        v = load_zin(a, grow, d);
        if (a.z0_out && !a.zin) a.z0_out[grow * D + d] = v;
        if (KIND == MNF_FLOW_RNVP) v *= noise_bern1(a.bern, grow, (uint32_t)d, D, 0.5f);
      }
      zs[rr][dd] = v;
    }
    // W1 chunk [32 d][H] -> wsm[0][h][dd] (transposed so stage-1 reads are conflict-free)
    for (int e = tid; e < FT_DC * H; e += FT_THREADS) {
      int dd = e / H, h = e - dd * H;
      int d = d0 + dd;
      float w = 0.f;
      if (d < D) {
        w = a.w1[(int64_t)d * H + h];
        if (a.m1) w *= a.m1[(int64_t)d * H + h];
      }
      wsm[0][h][dd] = w;
    }
    __syncthreads();
#pragma unroll 4
    for (int dd = 0; dd < FT_DC; ++dd) {
      float zv = zs[r][dd];
#pragma unroll
      for (int j = 0; j < FT_JMAX; ++j) {
        int h = g + 8 * j;
        if (h < H) acc[j] = fmaf(zv, wsm[0][h][dd], acc[j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < FT_JMAX; ++j) {
    int h = g + 8 * j;
    if (h < H) {
      float pre = acc[j] + a.b1[h];
      float hv = (KIND == MNF_FLOW_IAF) ? fmaxf(pre, 0.f) : tanhf(pre);
      hs[r][h] = hv;
      if (a.hid_save && row_ok) a.hid_save[row * H + h] = hv;
    }
  }
  __syncthreads();

  // ---- stage 2: (s,t) = h @ [Ws|Wt] + [bs|bt], combine, per-row log-det
  float ld = 0.f;
  for (int c0 = 0; c0 < D; c0 += FT_DC) {
    for (int e = tid; e < H * FT_DC; e += FT_THREADS) {
      int h = e >> 5, cc = e & 31;
      int c = c0 + cc;
      float vs = 0.f, vt = 0.f;
      if (c < D) {
        vs = a.ws[(int64_t)h * a.ld2 + c];
        vt = a.wt[(int64_t)h * a.ld2 + c];
        if (a.m2) {  // IAF: m2 [H,2D]; scale half at column c, shift half at D + c
          vs *= a.m2[(int64_t)h * a.ld2 + c];
          vt *= a.m2[(int64_t)h * a.ld2 + D + c];
        }
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
    if (row_ok) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int c = c0 + g + 8 * q;
        if (c < D) {
          float sv = s[q] + a.bs[c], tv = t[q] + a.bt[c];
          float zv = a.zin ? a.zin[row * D + c] : load_zin(a, row, c);
          if (KIND == MNF_FLOW_IAF) {
            float y = fmaf(zv, expf(sv), tv);
            a.zout[row * D + (a.parity ? D - 1 - c : c)] = y;
            ld += sv;
          } else {
            float m = noise_bern1(a.bern, row, (uint32_t)c, D, 0.5f);
            float gate = 1.f / (1.f + expf(-sv));
            // log(sigmoid(sv)) = -softplus(-sv), stable form
            float lg = sv >= 0.f ? -log1pf(expf(-sv)) : sv - log1pf(expf(sv));
            float x = ((1.f - m) * zv * gate + (1.f - gate) * tv) + m * zv;
            a.zout[row * D + c] = x;
            ld += (1.f - m) * lg;
          }
        }
      }
    }
    __syncthreads();
  }
  ld += __shfl_xor_sync(0xffffffffu, ld, 1);
  ld += __shfl_xor_sync(0xffffffffu, ld, 2);
  ld += __shfl_xor_sync(0xffffffffu, ld, 4);
  if (g == 0 && row_ok && a.ld_out) a.ld_out[row] = (a.ld_in ? a.ld_in[row] : 0.f) + ld;
}

// ------------------------------------------------------------------------------ planar
// scal[0] = k = (softplus(u.w) - 1 - u.w) / |w|^2 ; scal[1] = w . u_hat   (planar.py:23-26,33-34)
__global__ void planar_prep(int D, const float* __restrict__ w, const float* __restrict__ u, float* scal) {
  __shared__ double red[3][32];
  double uw = 0, ww = 0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    uw += (double)u[d] * w[d];
    ww += (double)w[d] * w[d];
  }
  uw = warp_sum_d(uw);
  ww = warp_sum_d(ww);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red[0][wid] = uw; red[1][wid] = ww; }
  __syncthreads();
  if (wid == 0) {
    uw = lane < (blockDim.x >> 5) ? red[0][lane] : 0;
    ww = lane < (blockDim.x >> 5) ? red[1][lane] : 0;
    uw = warp_sum_d(uw);
    ww = warp_sum_d(ww);
    if (lane == 0) { red[0][0] = uw; red[1][0] = ww; }
  }
  __syncthreads();
  float fuw = (float)red[0][0], fww = (float)red[1][0];
  float sp = fuw > 20.f ? fuw : log1pf(expf(fuw));
  float k = ((-1.f + sp) - fuw) / fww;
  // w . u_hat with u_hat = u + k w
  double wu = 0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) wu += (double)w[d] * (u[d] + k * w[d]);
  wu = warp_sum_d(wu);
  __syncthreads();
  if (lane == 0) red[2][wid] = wu;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[2][i];
    scal[0] = k;
    scal[1] = (float)t;
  }
}

struct PlanarArgs {
  int64_t B;
  int D;
  const float* zin;
  const float* q0_mean;
  const float* q0_log_var;
  NoiseDev eps;
  float* z0_out;
  float* zout;
  const float* ld_in;
  float* ld_out;
  const float *w, *b, *u;
  const float* scal;
};

// one warp per row: a = z.w + b (shuffle reduction), x = z + tanh(a) u_hat, ld = log|1 + (1-tanh^2) w.u_hat|
__global__ void __launch_bounds__(256) planar_step_fwd(PlanarArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= a.B) return;
  const int D = a.D;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) {
    float zv;
    if (a.zin) {
      zv = a.zin[row * D + d];
    } else {
      zv = a.q0_mean[d] + expf(0.5f * a.q0_log_var[d]) * noise_normal1(a.eps, row, 0u, (uint32_t)d, 1, D);
      if (a.z0_out) a.z0_out[row * D + d] = zv;
    }
    dot = fmaf(zv, a.w[d], dot);
  }
  dot = warp_sum(dot);
  const float th = tanhf(dot + a.b[0]);
  const float k = a.scal[0], wu = a.scal[1];
  for (int d = lane; d < D; d += 32) {
    float zv = a.zin ? a.zin[row * D + d]
                     : a.q0_mean[d] + expf(0.5f * a.q0_log_var[d]) * noise_normal1(a.eps, row, 0u, (uint32_t)d, 1, D);
    a.zout[row * D + d] = fmaf(th, a.u[d] + k * a.w[d], zv);
  }
  if (lane == 0 && a.ld_out) {
    float ld = logf(fabsf(1.f + (1.f - th * th) * wu));
    a.ld_out[row] = (a.ld_in ? a.ld_in[row] : 0.f) + ld;
  }
}

// T == 0: states[0] = z0 (sampled or copied), log_dets = 0
__global__ void sample_z0_kernel(int64_t B, int D, const float* zin, const float* q0_mean, const float* q0_log_var,
                                 NoiseDev eps, float* zout, float* ld_out) {
  int64_t total = B * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / D;
    int d = (int)(i - r * D);
    float v = zin ? zin[i] : q0_mean[d] + expf(0.5f * q0_log_var[d]) * noise_normal1(eps, r, 0u, (uint32_t)d, 1, D);
    zout[i] = v;
    if (ld_out && d == 0) ld_out[r] = 0.f;
  }
}

extern "C" int mnf_flow_forward(mnf_ctx* ctx, int64_t B, int64_t D, int32_t T, const mnf_flow_step* steps,
                                const float* z0, const float* q0_mean, const float* q0_log_var,
                                const mnf_noise* eps_z, float* states, float* log_dets, float* saved) {
  MNF_CHECK_ARG(ctx && states, "mnf_flow_forward: ctx/states is NULL");
  MNF_CHECK_ARG(B >= 0 && D >= 1 && D < (1 << 30) && T >= 0 && T <= 64, "mnf_flow_forward: bad B=%lld D=%lld T=%d",
                (long long)B, (long long)D, T);
  MNF_CHECK_ARG(T == 0 || steps, "mnf_flow_forward: steps is NULL");
  MNF_CHECK_ARG(z0 || (q0_mean && q0_log_var), "mnf_flow_forward: need z0 or (q0_mean, q0_log_var)");
  if (B == 0) return MNF_OK;
  NoiseDev eps = make_noise(eps_z);
  const int64_t BD = B * D;
  if (T == 0) {
    int grid = (int)(cdiv(BD, 256) < ctx->sm_count * 8 ? cdiv(BD, 256) : ctx->sm_count * 8);
    sample_z0_kernel<<<grid, 256, 0, ctx->stream>>>(B, (int)D, z0, q0_mean, q0_log_var, eps, states, log_dets);
    MNF_LAUNCHED(ctx);
    return MNF_OK;
  }
  // planar scalars live in the workspace: [T][2]
  float* scal = nullptr;
  bool any_planar = false;
  for (int k = 0; k < T; ++k) any_planar |= steps[k].kind == MNF_FLOW_PLANAR;
  if (any_planar) {
    scal = (float*)mnf_workspace(ctx, sizeof(float) * 2 * T);
    if (!scal) return MNF_ERR_CUDA;
  }
  int64_t saved_off = 0;  // saved is packed: step k starts after sum_{j<k} B*hidden_j
  for (int k = 0; k < T; ++k) {
    const mnf_flow_step& st = steps[k];
    const bool first = (k == 0);
    const float* zin = first ? z0 : states + (int64_t)k * BD;
    float* zout = states + (int64_t)(k + 1) * BD;
    float* z0_out = (first && !z0) ? states : nullptr;
    if (first && z0 && z0 != states) MNF_CUDA(cudaMemcpyAsync(states, z0, sizeof(float) * BD, cudaMemcpyDeviceToDevice, ctx->stream));
    if (st.kind == MNF_FLOW_IAF || st.kind == MNF_FLOW_RNVP) {
      MNF_CHECK_ARG(st.hidden >= 1 && st.hidden <= FT_HMAX, "flow step %d: hidden=%d not in [1,%d]", k, st.hidden, FT_HMAX);
      MNF_CHECK_ARG(st.w1 && st.b1 && st.ws && st.bs && st.wt && st.bt, "flow step %d: NULL weight pointer", k);
      FlowStepArgs a;
      a.B = B; a.D = (int)D; a.H = st.hidden; a.parity = st.parity;
      a.zin = zin; a.q0_mean = q0_mean; a.q0_log_var = q0_log_var; a.eps = eps;
      a.z0_out = z0_out; a.zout = zout;
      a.ld_in = first ? nullptr : log_dets; a.ld_out = log_dets;
      a.hid_save = saved ? saved + saved_off : nullptr;
      saved_off += B * st.hidden;
      a.w1 = st.w1; a.b1 = st.b1; a.ws = st.ws; a.bs = st.bs; a.wt = st.wt; a.bt = st.bt;
      a.m1 = st.m1; a.m2 = st.m2; a.ld2 = st.ld2;
      a.bern = make_noise(&st.bern);
      int grid = (int)cdiv(B, FT_ROWS);
      if (st.kind == MNF_FLOW_IAF)
        flow_mlp_step_fwd<MNF_FLOW_IAF><<<grid, FT_THREADS, 0, ctx->stream>>>(a);
      else
        flow_mlp_step_fwd<MNF_FLOW_RNVP><<<grid, FT_THREADS, 0, ctx->stream>>>(a);
      MNF_LAUNCHED(ctx);
    } else if (st.kind == MNF_FLOW_PLANAR) {
      MNF_CHECK_ARG(st.w1 && st.b1 && st.wt, "planar step %d: NULL pointer", k);
      planar_prep<<<1, 256, 0, ctx->stream>>>((int)D, st.w1, st.wt, scal + 2 * k);
      MNF_LAUNCHED(ctx);
      PlanarArgs a;
      a.B = B; a.D = (int)D; a.zin = zin; a.q0_mean = q0_mean; a.q0_log_var = q0_log_var; a.eps = eps;
      a.z0_out = z0_out; a.zout = zout; a.ld_in = first ? nullptr : log_dets; a.ld_out = log_dets;
      a.w = st.w1; a.b = st.b1; a.u = st.wt; a.scal = scal + 2 * k;
      planar_step_fwd<<<(int)cdiv(B, 8), 256, 0, ctx->stream>>>(a);
      MNF_LAUNCHED(ctx);
    } else {
      MNF_CHECK_ARG(false, "flow step %d: unknown kind %d", k, st.kind);
    }
  }
  return MNF_OK;
}
