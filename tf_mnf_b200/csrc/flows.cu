// Normalizing-flow steps as fused row-tile kernels (forward).
//   IAF   : maf.py:47-53 (MAF.inverse bound as IAF.forward, maf.py:63), MADE made.py:77-91,
//           MaskedDense made.py:28-30
//   RNVP  : rnvp.py:33-47
//   planar: planar.py:22-36
//   chain : flows/__init__.py:22-29;  z0 sampling: mnf_dense.py:93-96 / mnf_conv.py:105-108
// One launch per flow step (flows_tile.cuh); kl_div's single-row chains run as ONE launch
// (flow_chain_row1).  The D->H->2D coupling MLP runs out of shared memory (weights streamed
// in 32-wide chunks); the per-row log-det is reduced with warp shuffles.
#include "common.cuh"


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

#include "flows_tile.cuh"

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

// The same step for D % 4 == 0, D <= 4096: the row lives in registers (NV float4 per lane, WPR
// warps per row, coalesced 16-byte accesses), so z is read once, a sampled z0 costs one Philox
// call per 4 elements, and wide rows (D = 4096: 32 KB in + out per row) move at HBM speed.  All
// loads of a row are issued before the first use (the scalar kernel's load -> use -> load chain
// costs a memory round trip per 32 columns); at most 8 float4 per lane keep 32 warps per SM.
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <int NV, int WPR, bool SAMPLE>
__global__ void __launch_bounds__(256) planar_step_vec(PlanarArgs a) {
  __shared__ float part[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = wid % WPR;                                   // this warp's slice of the row
  const int64_t row = (int64_t)blockIdx.x * (8 / WPR) + wid / WPR;
  const bool row_ok = row < a.B;
  const int D = a.D;
  const float* __restrict__ zin = a.zin;
  const float* __restrict__ wv = a.w;
  const float* __restrict__ uv = a.u;
  float4 zv[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int d0 = 4 * ((j * WPR + sub) * 32 + lane);
    zv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_ok && d0 < D) {
      if (!SAMPLE) {
        zv[j] = ld4(zin + row * D + d0);
      } else {
        float n[4];
        if (a.eps.injected) {
          const float4 t = ld4(a.eps.injected + row * D + d0);
          n[0] = t.x; n[1] = t.y; n[2] = t.z; n[3] = t.w;
        } else {
          philox_normal4(a.eps, a.eps.row_offset + (uint64_t)row, 0u, (uint32_t)(d0 >> 2), n);
        }
        const float4 qm = ld4(a.q0_mean + d0), ql = ld4(a.q0_log_var + d0);
        zv[j].x = qm.x + expf(0.5f * ql.x) * n[0];
        zv[j].y = qm.y + expf(0.5f * ql.y) * n[1];
        zv[j].z = qm.z + expf(0.5f * ql.z) * n[2];
        zv[j].w = qm.w + expf(0.5f * ql.w) * n[3];
      }
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int d0 = 4 * ((j * WPR + sub) * 32 + lane);
    if (row_ok && d0 < D) {
      if (SAMPLE && a.z0_out) *reinterpret_cast<float4*>(a.z0_out + row * D + d0) = zv[j];
      const float4 w = ld4(wv + d0);
      dot = fmaf(zv[j].x, w.x, dot);
      dot = fmaf(zv[j].y, w.y, dot);
      dot = fmaf(zv[j].z, w.z, dot);
      dot = fmaf(zv[j].w, w.w, dot);
    }
  }
  dot = warp_sum(dot);
  if (WPR > 1) {
    if (lane == 0) part[wid] = dot;
    __syncthreads();
    dot = 0.f;
#pragma unroll
    for (int i = 0; i < WPR; ++i) dot += part[(wid / WPR) * WPR + i];  // fixed order: same sum in every warp
  }
  const float th = tanhf(dot + a.b[0]);
  const float k = a.scal[0], wu = a.scal[1];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int d0 = 4 * ((j * WPR + sub) * 32 + lane);
    if (row_ok && d0 < D) {
      const float4 w = ld4(wv + d0), u = ld4(uv + d0);
      float4 o;
      o.x = fmaf(th, u.x + k * w.x, zv[j].x);
      o.y = fmaf(th, u.y + k * w.y, zv[j].y);
      o.z = fmaf(th, u.z + k * w.z, zv[j].z);
      o.w = fmaf(th, u.w + k * w.w, zv[j].w);
      *reinterpret_cast<float4*>(a.zout + row * D + d0) = o;
    }
  }
  if (row_ok && sub == 0 && lane == 0 && a.ld_out) {
    float ld = logf(fabsf(1.f + (1.f - th * th) * wu));
    a.ld_out[row] = (a.ld_in ? a.ld_in[row] : 0.f) + ld;
  }
}
static bool planar_vec_ok(const PlanarArgs& a) {
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  return a.D % 4 == 0 && a.D <= 4096 && al(a.zin) && al(a.zout) && al(a.z0_out) && al(a.w) && al(a.u) &&
         al(a.q0_mean) && al(a.q0_log_var) && al(a.eps.injected);
}
template <int NV, int WPR>
static void planar_launch2(mnf_ctx* ctx, const PlanarArgs& a) {
  const int grid = (int)cdiv(a.B, 8 / WPR);
  if (a.zin) planar_step_vec<NV, WPR, false><<<grid, 256, 0, ctx->stream>>>(a);
  else planar_step_vec<NV, WPR, true><<<grid, 256, 0, ctx->stream>>>(a);
}
static void planar_launch(mnf_ctx* ctx, const PlanarArgs& a) {
  if (!planar_vec_ok(a)) { planar_step_fwd<<<(int)cdiv(a.B, 8), 256, 0, ctx->stream>>>(a); return; }
  const int nv = (a.D + 127) / 128;  // which variant depends on D only: any row split takes the same path
  if (nv <= 1) planar_launch2<1, 1>(ctx, a);
  else if (nv <= 2) planar_launch2<2, 1>(ctx, a);
  else if (nv <= 4) planar_launch2<4, 1>(ctx, a);
  else if (nv <= 8) planar_launch2<8, 1>(ctx, a);
  else if (nv <= 16) planar_launch2<8, 2>(ctx, a);
  else planar_launch2<8, 4>(ctx, a);
}

// Wide IAF steps on few rows (the VGG config's 4096-wide dense layer at batch 256): the row-tile
// kernel would run 8 CTAs through D/32 barrier-separated chunks twice (2.4 ms per step).  Here the two
// MLP contractions are plain GEMMs over the whole chip (split-K; tcgen05 in the tensor-core modes,
// with TMEM chains of <= 1024 terms) and two small kernels do the rest.
__global__ void iaf_hidden_kernel(int64_t n, int H, const float* __restrict__ b1, float* __restrict__ pre,
                                  float* __restrict__ hid_save) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = fmaxf(pre[i] + b1[(int)(i % H)], 0.f);
    pre[i] = v;
    if (hid_save) hid_save[i] = v;
  }
}
// y = z * exp(s) + t (reversed when parity), ld = sum_d s; a warp per row, st = [B][s(D) | t(D)] without biases
__global__ void __launch_bounds__(256) iaf_combine_kernel(int64_t B, int D, int parity, const float* __restrict__ zin,
                                                          const float* __restrict__ st, const float* __restrict__ bs,
                                                          const float* __restrict__ bt, float* __restrict__ zout,
                                                          const float* __restrict__ ld_in, float* __restrict__ ld_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* srow = st + row * 2 * D;
  float ld = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float s = srow[c] + bs[c], t = srow[D + c] + bt[c];
    zout[row * D + (parity ? D - 1 - c : c)] = fmaf(zin[row * D + c], expf(s), t);
    ld += s;
  }
  ld = warp_sum(ld);
  if (lane == 0 && ld_out) ld_out[row] = (ld_in ? ld_in[row] : 0.f) + ld;
}

// T == 0: states[0] = z0 (sampled or copied), log_dets = 0.  One thread per 4 consecutive
// elements of a row: one Philox call yields the four normals.
__global__ void sample_z0_kernel(int64_t B, int D, const float* zin, const float* q0_mean, const float* q0_log_var,
                                 NoiseDev eps, float* zout, float* ld_out) {
  const int D4 = (D + 3) >> 2;
  const int64_t total = B * D4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D4;
    const int g4 = (int)(i - r * D4), d0 = g4 * 4;
    float nrm[4] = {0.f, 0.f, 0.f, 0.f};
    if (!zin && !eps.injected) philox_normal4(eps, eps.row_offset + (uint64_t)r, 0u, (uint32_t)g4, nrm);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = d0 + e;
      if (d < D) {
        float v;
        if (zin) v = zin[r * D + d];
        else {
          const float n = eps.injected ? eps.injected[r * D + d] : nrm[e];
          v = q0_mean[d] + expf(0.5f * q0_log_var[d]) * n;
        }
        zout[r * D + d] = v;
      }
    }
    if (ld_out && g4 == 0) ld_out[r] = 0.f;
  }
}

// ------------------------------------------------------------- single-row chain (kl_div)
// kl_div calls sample_z(1) and flow_r.forward on ONE row (mnf_dense.py:105,135): the row-tile
// kernel would run a single CTA through 2*D/32 barrier-separated chunks.  Here one 1024-thread
// CTA runs the WHOLE chain: stage 1 splits the D-reduction over 16 thread slices, stage 2
// gives every thread output columns; z ping-pongs in shared memory between steps.
#define C1_THREADS 1024
#define C1_DMAX 4096
#define C1_TMAX 8

struct Chain1Step {
  int kind, parity, H;
  const float *w1, *b1, *ws, *bs, *wt, *bt, *m1, *m2;
  int64_t ld2;
  NoiseDev bern;
};
struct Chain1Args {
  int D, T;
  const float* zin;
  const float* q0_mean;
  const float* q0_log_var;
  NoiseDev eps;
  float* states;    // [T+1, D]
  float* ld_out;    // [1] or nullptr
  float* hid_save;  // packed [sum H] or nullptr
  Chain1Step st[C1_TMAX];
};

__global__ void __launch_bounds__(C1_THREADS) flow_chain_row1(Chain1Args a) {
  __shared__ float zbuf[2][C1_DMAX];
  __shared__ float part[16][FT_HMAX + 1];
  __shared__ float hs[FT_HMAX];
  __shared__ float red[32];
  const int tid = threadIdx.x, D = a.D;
  for (int d = tid; d < D; d += C1_THREADS) {
    float v = a.zin ? a.zin[d]
                    : a.q0_mean[d] + expf(0.5f * a.q0_log_var[d]) * noise_normal1(a.eps, 0, 0u, (uint32_t)d, 1, D);
    zbuf[0][d] = v;
    a.states[d] = v;
  }
  __syncthreads();
  float ld_tot = 0.f;
  int hoff = 0;
  for (int k = 0; k < a.T; ++k) {
    const Chain1Step& s = a.st[k];
    const float* zc = zbuf[k & 1];
    float* zn = zbuf[(k + 1) & 1];
    const int H = s.H;
    // stage 1
    {
      const int j = tid & 63, sl = tid >> 6;
      float acc = 0.f;
      if (j < H) {
        const int per = (D + 15) / 16;
        const int d0 = sl * per, d1 = min(D, d0 + per);
        for (int d = d0; d < d1; ++d) {
          float w = __ldg(s.w1 + (int64_t)d * H + j);
          if (s.m1) w *= __ldg(s.m1 + (int64_t)d * H + j);
          float zv = zc[d];
          if (s.kind == MNF_FLOW_RNVP) zv *= noise_bern1(s.bern, 0, (uint32_t)d, D, 0.5f);
          acc = fmaf(zv, w, acc);
        }
        part[sl][j] = acc;
      }
    }
    __syncthreads();
    if (tid < H) {
      float pre = s.b1[tid];
      for (int q = 0; q < 16; ++q) pre += part[q][tid];
      float hv = (s.kind == MNF_FLOW_IAF) ? fmaxf(pre, 0.f) : tanhf(pre);
      hs[tid] = hv;
      if (a.hid_save) a.hid_save[hoff + tid] = hv;
    }
    __syncthreads();
    // stage 2
    float ld = 0.f;
    for (int c = tid; c < D; c += C1_THREADS) {
      float sv = s.bs[c], tv = s.bt[c];
      for (int h = 0; h < H; ++h) {
        float ws = __ldg(s.ws + (int64_t)h * s.ld2 + c), wt = __ldg(s.wt + (int64_t)h * s.ld2 + c);
        if (s.m2) {
          ws *= __ldg(s.m2 + (int64_t)h * s.ld2 + c);
          wt *= __ldg(s.m2 + (int64_t)h * s.ld2 + D + c);
        }
        sv = fmaf(hs[h], ws, sv);
        tv = fmaf(hs[h], wt, tv);
      }
      float zv = zc[c];
      if (s.kind == MNF_FLOW_IAF) {
        float y = fmaf(zv, expf(sv), tv);
        int o = s.parity ? D - 1 - c : c;
        zn[o] = y;
        a.states[(int64_t)(k + 1) * D + o] = y;
        ld += sv;
      } else {
        float m = noise_bern1(s.bern, 0, (uint32_t)c, D, 0.5f);
        float gate = 1.f / (1.f + expf(-sv));
        float lg = sv >= 0.f ? -log1pf(expf(-sv)) : sv - log1pf(expf(sv));
        float x = ((1.f - m) * zv * gate + (1.f - gate) * tv) + m * zv;
        zn[c] = x;
        a.states[(int64_t)(k + 1) * D + c] = x;
        ld += (1.f - m) * lg;
      }
    }
    ld = warp_sum(ld);
    if ((tid & 31) == 0) red[tid >> 5] = ld;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int q = 0; q < C1_THREADS / 32; ++q) t += red[q];
      ld_tot += t;
    }
    hoff += H;
    __syncthreads();
  }
  if (tid == 0 && a.ld_out) a.ld_out[0] = ld_tot;
}

// -------------------------------------------------- single-row chain on a thread-block cluster
// The one-CTA kernel above pulls every step's ~(6*D*H) weight words through a single SM and is
// latency bound (65 us per step at D = 800).  Here a cluster of CL_N CTAs splits each step:
// stage 1 reduces a D-slice per CTA (partials exchanged through a tiny global scratch + a cluster
// barrier), stage 2 gives each CTA a column slice; the new state goes to `states` (which the
// reference returns anyway) and is re-read after the next cluster barrier.
#define CL_N 8
#define CL_THREADS 256
#define CL_MSK 512

struct ChainClArgs {
  Chain1Args c;
  float* part_h;   // [CL_N][FT_HMAX] stage-1 partial sums (re-used every step)
  float* part_ld;  // [CL_N]
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__global__ void __cluster_dims__(CL_N, 1, 1) __launch_bounds__(CL_THREADS) flow_chain_row1_cluster(ChainClArgs ca) {
  const Chain1Args& a = ca.c;
  __shared__ float part[4][FT_HMAX + 1];
  __shared__ float hs[FT_HMAX];
  __shared__ float red[CL_THREADS / 32];
  __shared__ float msk[CL_MSK];  // RNVP mask of this CTA's columns for the current step (one Philox per thread, once)
  const int tid = threadIdx.x, D = a.D;
  const int rank = (int)cluster_rank();
  // column / reduction slice of this CTA (multiples of 4 so one Philox call serves 4 draws)
  const int per = ((D + CL_N - 1) / CL_N + 3) & ~3;
  const int lo = min(D, rank * per), hi = min(D, lo + per);
  // states[0]
  for (int d = lo + tid; d < hi; d += CL_THREADS)
    a.states[d] = a.zin ? a.zin[d]
                        : a.q0_mean[d] + expf(0.5f * a.q0_log_var[d]) * noise_normal1(a.eps, 0, 0u, (uint32_t)d, 1, D);
  __threadfence();
  cluster_sync_all();
  float ld_tot = 0.f;
  int hoff = 0;
  for (int k = 0; k < a.T; ++k) {
    const Chain1Step& s = a.st[k];
    const float* zc = a.states + (int64_t)k * D;
    float* zn = a.states + (int64_t)(k + 1) * D;
    const int H = s.H;
    const bool use_msk = s.kind == MNF_FLOW_RNVP && hi - lo <= CL_MSK;
    if (use_msk) {
      for (int i = tid; i < hi - lo; i += CL_THREADS) msk[i] = noise_bern1(s.bern, 0, (uint32_t)(lo + i), D, 0.5f);
      __syncthreads();
    }
    // ---- stage 1: partial h over d in [lo, hi); thread = (hidden j, one of 4 d-sub-slices)
    {
      const int j = tid & 63, sl = tid >> 6;
      float acc = 0.f;
      if (j < H) {
        const int n = hi - lo, sub = ((n + 3) / 4 + 3) & ~3;
        const int d0 = lo + sl * sub, d1 = min(hi, d0 + sub);
        for (int dq = d0; dq < d1; dq += 4) {
          float m4[4] = {1.f, 1.f, 1.f, 1.f};
          if (use_msk) {
#pragma unroll
            for (int e = 0; e < 4; ++e) m4[e] = dq + e < hi ? msk[dq + e - lo] : 0.f;
          } else if (s.kind == MNF_FLOW_RNVP) {
            noise_bern4(s.bern, 0, (uint32_t)(dq >> 2), D, 0.5f, m4);
          }
          float w[4], zv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int d = dq + e;
            const bool ok = d < d1;
            w[e] = ok ? __ldg(s.w1 + (int64_t)d * H + j) : 0.f;
            if (ok && s.m1) w[e] *= __ldg(s.m1 + (int64_t)d * H + j);
            zv[e] = ok ? zc[d] : 0.f;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) acc = fmaf(zv[e] * m4[e], w[e], acc);
        }
        part[sl][j] = acc;
      }
    }
    __syncthreads();
    if (tid < FT_HMAX) ca.part_h[rank * FT_HMAX + tid] = tid < H ? part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid] : 0.f;
    __threadfence();
    cluster_sync_all();
    if (tid < H) {
      float pre = s.b1[tid];
#pragma unroll
      for (int q = 0; q < CL_N; ++q) pre += __ldcg(ca.part_h + q * FT_HMAX + tid);
      const float hv = (s.kind == MNF_FLOW_IAF) ? fmaxf(pre, 0.f) : tanhf(pre);
      hs[tid] = hv;
      if (a.hid_save && rank == 0) a.hid_save[hoff + tid] = hv;
    }
    __syncthreads();
    // ---- stage 2: this CTA's columns, two threads per column (each half of the hidden units: the
    //      dependent batches of weight loads are the latency of this kernel), combined by a shuffle
    float ld = 0.f;
    const int hsel = tid & 1, Hh = (H + 1) >> 1;
    const int hb0 = hsel * Hh, hb1 = min(H, hb0 + Hh);
    const int ncol_round = (hi - lo + (CL_THREADS >> 1) - 1) / (CL_THREADS >> 1) * (CL_THREADS >> 1);  // same trip count for all
    for (int cp = tid >> 1; cp < ncol_round; cp += CL_THREADS >> 1) {
      const int c = lo + cp;
      const bool cok = c < hi;  // both threads of a pair agree; idle pairs still take part in the shuffle
      float sv = 0.f, tv = 0.f;
      if (cok) {
        for (int h0 = hb0; h0 < hb1; h0 += 13) {
          float ws[13], wt[13];
#pragma unroll
          for (int q = 0; q < 13; ++q) {
            const int h = h0 + q;
            ws[q] = wt[q] = 0.f;
            if (h < hb1) {
              ws[q] = __ldg(s.ws + (int64_t)h * s.ld2 + c);
              wt[q] = __ldg(s.wt + (int64_t)h * s.ld2 + c);
              if (s.m2) {
                ws[q] *= __ldg(s.m2 + (int64_t)h * s.ld2 + c);
                wt[q] *= __ldg(s.m2 + (int64_t)h * s.ld2 + D + c);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 13; ++q)
            if (h0 + q < hb1) {
              sv = fmaf(hs[h0 + q], ws[q], sv);
              tv = fmaf(hs[h0 + q], wt[q], tv);
            }
        }
      }
      sv += __shfl_xor_sync(0xffffffffu, sv, 1);
      tv += __shfl_xor_sync(0xffffffffu, tv, 1);
      if (!cok || hsel) continue;
      sv += s.bs[c];
      tv += s.bt[c];
      const float zv = zc[c];
      if (s.kind == MNF_FLOW_IAF) {
        zn[s.parity ? D - 1 - c : c] = fmaf(zv, expf(sv), tv);
        ld += sv;
      } else {
        const float m = use_msk ? msk[cp] : noise_bern1(s.bern, 0, (uint32_t)c, D, 0.5f);
        const float gate = 1.f / (1.f + expf(-sv));
        const float lg = sv >= 0.f ? -log1pf(expf(-sv)) : sv - log1pf(expf(sv));
        zn[c] = ((1.f - m) * zv * gate + (1.f - gate) * tv) + m * zv;
        ld += (1.f - m) * lg;
      }
    }
    ld = warp_sum(ld);
    if ((tid & 31) == 0) red[tid >> 5] = ld;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int q = 0; q < CL_THREADS / 32; ++q) t += red[q];
      ld_tot += t;
    }
    hoff += H;
    __threadfence();
    cluster_sync_all();  // new state visible to every CTA; part_h / hs / red free for the next step
  }
  if (tid == 0) ca.part_ld[rank] = ld_tot;
  __threadfence();
  cluster_sync_all();
  if (rank == 0 && tid == 0 && a.ld_out) {
    float t = 0.f;
    for (int q = 0; q < CL_N; ++q) t += __ldcg(ca.part_ld + q);
    a.ld_out[0] = t;
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
  bool any_planar = false, all_mlp = true, any_rnvp = false;
  for (int k = 0; k < T; ++k) {
    any_planar |= steps[k].kind == MNF_FLOW_PLANAR;
    any_rnvp |= steps[k].kind == MNF_FLOW_RNVP;
    all_mlp &= (steps[k].kind == MNF_FLOW_IAF || steps[k].kind == MNF_FLOW_RNVP) && steps[k].hidden >= 1 &&
               steps[k].hidden <= FT_HMAX && steps[k].w1 && steps[k].b1 && steps[k].ws && steps[k].bs && steps[k].wt &&
               steps[k].bt;
  }
  bool any_generic = false;
  for (int k = 0; k < T; ++k) any_generic |= mnf_flow_step_is_generic(steps[k]);
  if (B == 1 && all_mlp && !any_generic && T <= C1_TMAX && D <= C1_DMAX) {  // kl_div path: whole chain in one CTA
    Chain1Args a;
    a.D = (int)D; a.T = T; a.zin = z0; a.q0_mean = q0_mean; a.q0_log_var = q0_log_var; a.eps = eps;
    a.states = states; a.ld_out = log_dets; a.hid_save = saved;
    for (int k = 0; k < T; ++k) {
      Chain1Step& s = a.st[k];
      s.kind = steps[k].kind; s.parity = steps[k].parity; s.H = steps[k].hidden;
      s.w1 = steps[k].w1; s.b1 = steps[k].b1; s.ws = steps[k].ws; s.bs = steps[k].bs; s.wt = steps[k].wt;
      s.bt = steps[k].bt; s.m1 = steps[k].m1; s.m2 = steps[k].m2; s.ld2 = steps[k].ld2;
      s.bern = make_noise(&steps[k].bern);
    }
    const bool probed = mnf_probe_begin(ctx, MNF_PROBE_FLOW_STEP);
    if (D >= 16) {  // split the chain over a cluster of CL_N CTAs (scratch: [CL_N][64] + [CL_N] floats)
      ChainClArgs ca;
      ca.c = a;
      float* scratch = (float*)mnf_workspace(ctx, sizeof(float) * (CL_N * FT_HMAX + CL_N + 64));
      if (!scratch) return MNF_ERR_CUDA;
      ca.part_h = scratch;
      ca.part_ld = scratch + CL_N * FT_HMAX;
      flow_chain_row1_cluster<<<CL_N, CL_THREADS, 0, ctx->stream>>>(ca);
    } else {
      flow_chain_row1<<<1, C1_THREADS, 0, ctx->stream>>>(a);
    }
    mnf_probe_end(ctx, probed);
    MNF_LAUNCHED(ctx);
    return MNF_OK;
  }
  // planar scalars live in the workspace: [T][2]
  // workspace: planar scalars [T][2] (padded to 256 B), then the packed split weights of one
  // tensor-core flow step (re-used step after step: launches are ordered on the stream)
  float* scal = nullptr;
  float* pack_ws = nullptr;
  // Tensor-core flow steps (flows_umma.cu).  Its stage-1 contraction accumulates D terms in ONE TMEM
  // accumulator and the tensor core's fp32 accumulation is not round-to-nearest: measured against the
  // float64 oracle the states stay at 2e-6 but log-dets / hidden activations reach 1.1e-5 at D = 2048
  // and 2.2e-5 at D = 4096 (FFMA: 4e-7), and the hidden activations cross the 1e-5 parity bar at
  // D = 896.  Wider rows take the GEMM formulation below (chains of <= 1024 terms) or the FFMA tiles.
  const int tc_dmax = 800;
  const bool use_tc = ctx->math_mode != MNF_MATH_FP32 && B >= 32 && D <= tc_dmax;
  // wide rows on few row tiles: the MLP contractions as whole-chip GEMMs (iaf_combine_kernel above)
  const bool use_gemm = !use_tc && D >= 1024 && B >= 2 && cdiv(B, FT_ROWS) * 2 <= ctx->sm_count;
  float* gemm_ws = nullptr;
  // RNVP steps of many rows: flows_f16.cu (needs the max |z| of every 128-row tile: two ping-pong arrays)
  bool any_f16 = false;
  for (int k = 0; k < T; ++k) any_f16 |= use_tc && mnf_flow_f16_ok(ctx, B, D, steps[k]);
  any_f16 = any_f16 && ((reinterpret_cast<uintptr_t>(states) | (z0 ? reinterpret_cast<uintptr_t>(z0) : 0)) & 15) == 0;
  float *tmax_cur = nullptr, *tmax_nxt = nullptr;
  uint8_t* pack16 = nullptr;
  bool tmax_valid = false;
  {
    const size_t scal_floats = ((size_t)2 * T + 63) & ~(size_t)63;
    const size_t pack_floats = use_tc ? mnf_flow_umma_pack_floats((int)D) : 0;
    const size_t gemm_floats = use_gemm ? (size_t)B * (64 + 2 * (size_t)D) : 0;
    const size_t tmax_floats = any_f16 ? ((mnf_flow_f16_tiles(B) + 63) & ~(size_t)63) : 0;
    const size_t pack16_floats = any_f16 ? (mnf_flow_f16_pack_bytes((int)D) + 3) / 4 : 0;
    if (any_planar || use_tc || use_gemm) {
      scal = (float*)mnf_workspace(ctx, sizeof(float) * (scal_floats + pack_floats + gemm_floats + 2 * tmax_floats + pack16_floats));
      if (!scal) return MNF_ERR_CUDA;
      pack_ws = scal + scal_floats;
      gemm_ws = pack_ws + pack_floats;
      tmax_cur = gemm_ws + gemm_floats;
      tmax_nxt = tmax_cur + tmax_floats;
      pack16 = (uint8_t*)(tmax_nxt + tmax_floats);
    }
  }
  int64_t saved_off = 0;  // saved is packed: step k starts after sum_{j<k} B*hidden_j
  for (int k = 0; k < T; ++k) {
    const mnf_flow_step& st = steps[k];
    const bool first = (k == 0);
    const float* zin = first ? z0 : states + (int64_t)k * BD;
    float* zout = states + (int64_t)(k + 1) * BD;
    float* z0_out = (first && !z0) ? states : nullptr;
    const bool had_tmax = tmax_valid;  // zin's tile maxima are known (written by the f16 step before this one)
    tmax_valid = false;
    if (first && z0 && z0 != states) MNF_CUDA(cudaMemcpyAsync(states, z0, sizeof(float) * BD, cudaMemcpyDeviceToDevice, ctx->stream));
    if (mnf_flow_step_is_generic(st)) {  // several hidden layers / another activation: the general kernels
      const float* zg = zin;
      if (!zg) {  // sample z0 for all rows first (same Philox counters as the fused draw)
        const int64_t n4 = B * ((D + 3) / 4);
        int g0 = (int)(cdiv(n4, 256) < ctx->sm_count * 8 ? cdiv(n4, 256) : ctx->sm_count * 8);
        sample_z0_kernel<<<g0, 256, 0, ctx->stream>>>(B, (int)D, nullptr, q0_mean, q0_log_var, eps, states, nullptr);
        MNF_LAUNCHED(ctx);
        zg = states;
      }
      int rc = mnf_flow_generic_forward(ctx, B, D, st, zg, zout, first ? nullptr : log_dets, log_dets,
                                        saved ? saved + saved_off : nullptr);
      if (rc) return rc;
      saved_off += B * mnf_flow_step_hidden_total(st);
      continue;
    }
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
      if (any_f16 && mnf_flow_f16_ok(ctx, B, D, st)) {  // tcgen05 kind::f16 path (flows_f16.cu)
        const float* zin16 = zin;
        if (!zin16) {  // z0 for all rows (same Philox counters as the fused draw) + its tile maxima
          int rc = mnf_flow_f16_sample_z0(ctx, B, (int)D, q0_mean, q0_log_var, eps, states, tmax_cur);
          if (rc) return rc;
          zin16 = states;
        } else if (!had_tmax) {
          int rc = mnf_flow_f16_absmax(ctx, B, (int)D, zin16, tmax_cur);
          if (rc) return rc;
        }
        int rc = mnf_flow_f16_step(ctx, B, (int)D, zin16, zout, a.ld_in, a.ld_out, a.hid_save, st, a.bern, tmax_cur, tmax_nxt, pack16);
        if (rc) return rc;
        float* t = tmax_cur; tmax_cur = tmax_nxt; tmax_nxt = t;
        tmax_valid = true;
        continue;
      }
      if (use_tc && (z0_out || zin)) {  // tcgen05 path (3xTF32): same arithmetic contract, fp32 parity
        const float* zin_tc = zin;
        float* z0_tc = z0_out;
        if (!zin) {
          // sample z0 for ALL rows with a full-chip elementwise kernel (same Philox counters) instead
          // of inside the 80 row-tile CTAs' producer warps, where it sits on the critical path
          const int64_t n4 = B * ((D + 3) / 4);
          int g0 = (int)(cdiv(n4, 256) < ctx->sm_count * 8 ? cdiv(n4, 256) : ctx->sm_count * 8);
          sample_z0_kernel<<<g0, 256, 0, ctx->stream>>>(B, (int)D, nullptr, q0_mean, q0_log_var, eps, states, nullptr);
          MNF_LAUNCHED(ctx);
          zin_tc = states;
          z0_tc = nullptr;
        }
        int rc = mnf_flow_umma_step(ctx, B, (int)D, st.hidden, st.kind, st.parity, zin_tc, q0_mean, q0_log_var, eps, z0_tc,
                                    zout, a.ld_in, a.ld_out, a.hid_save, st, a.bern, pack_ws);
        if (rc) return rc;
        continue;
      }
      if (use_gemm && st.kind == MNF_FLOW_IAF && !st.m1 && !st.m2 && st.wt == st.ws + D && st.ld2 == 2 * D) {
        const float* zg = zin;
        if (!zg) {  // sample z0 for all rows first (same Philox counters as the fused draw)
          const int64_t n4 = B * ((D + 3) / 4);
          int g0 = (int)(cdiv(n4, 256) < ctx->sm_count * 8 ? cdiv(n4, 256) : ctx->sm_count * 8);
          sample_z0_kernel<<<g0, 256, 0, ctx->stream>>>(B, (int)D, nullptr, q0_mean, q0_log_var, eps, states, nullptr);
          MNF_LAUNCHED(ctx);
          zg = states;
        }
        const int H = st.hidden;
        float* hbuf = gemm_ws;                   // [B, H]
        float* stbuf = gemm_ws + (size_t)B * 64;  // [B, 2D]
        int rc = mnf_gemm_any(ctx, (int)B, H, (int)D, zg, D, 1, st.w1, H, 1, hbuf, H, false);
        if (rc) return rc;
        const int64_t nh = B * H;
        iaf_hidden_kernel<<<(int)cdiv(nh, 256), 256, 0, ctx->stream>>>(nh, H, st.b1, hbuf, a.hid_save);
        MNF_LAUNCHED(ctx);
        rc = mnf_gemm_any(ctx, (int)B, (int)(2 * D), H, hbuf, H, 1, st.ws, 2 * D, 1, stbuf, 2 * D, false);
        if (rc) return rc;
        mnf_note_path("flow_gemm");
        iaf_combine_kernel<<<(int)cdiv(B, 8), 256, 0, ctx->stream>>>(B, (int)D, st.parity, zg, stbuf, st.bs, st.bt, zout,
                                                                   a.ld_in, a.ld_out);
        MNF_LAUNCHED(ctx);
        continue;
      }
      int grid = (int)cdiv(B, FT_ROWS);
      mnf_note_path("flow_mlp");
      if (st.kind == MNF_FLOW_IAF)
        flow_mlp_step_fwd<MNF_FLOW_IAF><<<grid, FT_THREADS, 0, ctx->stream>>>(a);
      else
        flow_mlp_step_fwd<MNF_FLOW_RNVP><<<grid, FT_THREADS, 0, ctx->stream>>>(a);
      MNF_LAUNCHED(ctx);
    } else if (st.kind == MNF_FLOW_PLANAR) {
      MNF_CHECK_ARG(st.w1 && st.b1 && st.wt, "planar step %d: NULL pointer", k);
      planar_prep<<<1, 1024, 0, ctx->stream>>>((int)D, st.w1, st.wt, scal + 2 * k);
      MNF_LAUNCHED(ctx);
      PlanarArgs a;
      a.B = B; a.D = (int)D; a.zin = zin; a.q0_mean = q0_mean; a.q0_log_var = q0_log_var; a.eps = eps;
      a.z0_out = z0_out; a.zout = zout; a.ld_in = first ? nullptr : log_dets; a.ld_out = log_dets;
      a.w = st.w1; a.b = st.b1; a.u = st.wt; a.scal = scal + 2 * k;
      planar_launch(ctx, a);
      MNF_LAUNCHED(ctx);
    } else {
      MNF_CHECK_ARG(false, "flow step %d: unknown kind %d", k, st.kind);
    }
  }
  return MNF_OK;
}
