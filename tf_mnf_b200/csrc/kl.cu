// MNFDense.kl_div (mnf_dense.py:104-157) and MNFConv2D.kl_div (mnf_conv.py:116-180).
//
// Forward: ONE launch.  Every CTA streams its slice of (mean_W, log_std_W) once with
// coalesced loads, accumulating the weight-KL sum and the two weighted reductions
// (mean_w, var_w: mnf_dense.py:139-140 / mnf_conv.py:154-155); the last CTA to finish
// (atomic ticket) folds the partials in a fixed order and evaluates the scalar tail
// (tanh / linear `a`, r-Gaussian term over the flow_r states, bias KL, q entropy).
// Backward: the same pass with the `aux` side-outputs, then a scalar kernel and one
// element-wise pass over the weights (SURVEY.md App. A-4/A-5 derivatives).
#include "common.cuh"

#define KL_THREADS 256
#define KL_ROWS 32  // rows per CTA in the dense pass
#define LOG2PI_F 1.8378770664093453f

struct KlDev {
  int conv, use_z, r_all, n_states;
  int rows, cols, D;
  const float *Wm, *Ls, *bm, *lvb, *p, *pb, *q0lv, *r0m, *r0lv, *cap, *z, *ldq, *rst, *ldr;
  NoiseDev eps_a, eps_b;
  // workspace
  double* partS;    // [nblk]
  float* part_mw;   // dense: [nchunk][cols]
  float* part_vw;
  float* mw;        // [L]  (conv: written directly)
  float* vw;
  float* aux;       // [L + 8]: a_j, then abar, mu_b, var_b
  float* dmw;       // [L] backward: dKL/d mean_w
  float* dvw;       // [L] backward: dKL/d var_w
  unsigned int* counter;
  int nchunk;
  float* kl_out;
};

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  double t = 0;
  if (wid == 0) {
    t = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    t = warp_sum_d(t);
    if (lane == 0) sh[0] = t;
  }
  __syncthreads();
  t = sh[0];
  __syncthreads();
  return t;
}

__device__ void kl_finish(const KlDev& k, int nblk, double* sh) {
  const int tid = threadIdx.x;
  const int L = k.conv ? k.rows : k.cols;  // length of the `a` vector
  // 1. weight KL
  double s = 0;
  for (int i = tid; i < nblk; i += blockDim.x) s += k.partS[i];
  const double klw = 0.5 * block_sum_d(s, sh);
  // 2. bias KL (mnf_dense.py:118-124 / mnf_conv.py:134-140)
  const double pb = k.pb[0], epb = exp(pb);
  double sb = 0;
  for (int j = tid; j < k.cols; j += blockDim.x) {
    double b = k.bm[j];
    if (k.conv && k.use_z) b *= k.z[j];
    double lv = k.lvb[j];
    sb += pb - lv + (exp(lv) + b * b) / epb - 1.0;
  }
  const double klb = 0.5 * block_sum_d(sb, sh);
  if (!k.use_z) {
    if (tid == 0) k.kl_out[0] = (float)(klw + klb);
    return;
  }
  // 3. log q (mnf_dense.py:126-130)
  double sq = 0;
  for (int d = tid; d < k.D; d += blockDim.x) sq += (double)LOG2PI_F + k.q0lv[d] + 1.0;
  const double logq = -(double)k.ldq[0] - 0.5 * block_sum_d(sq, sh);
  // 4. conv bias contribution to `a` (mnf_conv.py:163-165)
  double mu_b = 0, var_b = 0, add_b = 0;
  if (k.conv) {
    double m = 0, v = 0;
    for (int f = tid; f < k.cols; f += blockDim.x) {
      double c = k.cap[f];
      m += (double)k.bm[f] * k.z[f] * c;
      v += exp((double)k.lvb[f]) * c * c;
    }
    mu_b = block_sum_d(m, sh);
    var_b = block_sum_d(v, sh);
    float eb = noise_normal1(k.eps_b, 0, 0u, 0u, 1, 1);
    add_b = mu_b + sqrt(var_b) * (double)eb;
  }
  // 5. a, abar
  double sa = 0;
  for (int j = tid; j < L; j += blockDim.x) {
    float mw, vw;
    if (k.conv) {
      mw = k.mw[j];
      vw = k.vw[j];
    } else {
      mw = 0.f;
      vw = 0.f;
      for (int c = 0; c < k.nchunk; ++c) {
        mw += k.part_mw[(size_t)c * k.cols + j];
        vw += k.part_vw[(size_t)c * k.cols + j];
      }
      k.mw[j] = mw;
      k.vw[j] = vw;
    }
    float ea = noise_normal1(k.eps_a, 0, 0u, (uint32_t)j, 1, L);
    float pre = mw + sqrtf(vw) * ea;
    float a = k.conv ? (float)(pre + add_b) : tanhf(pre);
    k.aux[j] = a;
    sa += a;
  }
  const double abar = block_sum_d(sa, sh) / (double)L;
  // 6. Gaussian r-term over the flow_r states (mnf_dense.py:151-155; App. B-3)
  double sl = 0;
  const int s0 = k.r_all ? 0 : k.n_states - 1;
  for (int d = tid; d < k.D; d += blockDim.x) {
    double mr = abar * k.r0m[d], lv = abar * k.r0lv[d], e = exp(lv);
    for (int st = s0; st < k.n_states; ++st) {
      double diff = (double)k.rst[(size_t)st * k.D + d] - mr;
      sl += -e * diff * diff - (double)LOG2PI_F + lv;
    }
  }
  const double logr = (double)k.ldr[0] + 0.5 * block_sum_d(sl, sh);
  if (tid == 0) {
    k.kl_out[0] = (float)(klw + klb - logr + logq);
    k.aux[L + 0] = (float)abar;
    k.aux[L + 1] = (float)mu_b;
    k.aux[L + 2] = (float)var_b;
  }
}

__global__ void __launch_bounds__(KL_THREADS) kl_fwd_kernel(KlDev k) {
  __shared__ double sh[32];
  __shared__ bool last;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nblk = gridDim.x * gridDim.y;
  const int blk = blockIdx.y * gridDim.x + blockIdx.x;
  double s = 0;
  if (!k.conv) {
    // CTA = KL_ROWS rows x KL_THREADS columns; thread owns one column
    const int j = blockIdx.y * KL_THREADS + tid;
    const int i0 = blockIdx.x * KL_ROWS;
    const int i1 = min(i0 + KL_ROWS, k.rows);
    float mw = 0.f, vw = 0.f;
    if (j < k.cols) {
      for (int i = i0; i < i1; ++i) {
        float ls = __ldg(k.Ls + (size_t)i * k.cols + j), wm = __ldg(k.Wm + (size_t)i * k.cols + j);
        float e = expf(ls), vt = e * e;
        float zi = k.use_z ? k.z[i] : 1.f;
        float mt = zi * wm;
        float pi = k.p[i];
        s += (double)(pi - 2.f * ls + (vt + mt * mt) / expf(pi) - 1.f);
        if (k.use_z) {
          float c = k.cap[i];
          mw = fmaf(mt, c, mw);
          vw = fmaf(vt, c * c, vw);
        }
      }
      if (k.use_z) {
        k.part_mw[(size_t)blockIdx.x * k.cols + j] = mw;
        k.part_vw[(size_t)blockIdx.x * k.cols + j] = vw;
      }
    }
  } else {
    // warp per row; lanes stride over the F columns
    const int i = blk * (KL_THREADS / 32) + wid;
    if (i < k.rows) {
      float mw = 0.f, vw = 0.f;
      float pi = k.p[i], epi = expf(pi);
      for (int f = lane; f < k.cols; f += 32) {
        float ls = __ldg(k.Ls + (size_t)i * k.cols + f), wm = __ldg(k.Wm + (size_t)i * k.cols + f);
        float e = expf(ls), vt = e * e;
        float zf = k.use_z ? k.z[f] : 1.f;
        float mt = zf * wm;
        s += (double)(pi - ls + (vt + mt * mt) / epi - 1.f);
        if (k.use_z) {
          float c = k.cap[f];
          mw = fmaf(mt, c, mw);
          vw = fmaf(vt, c * c, vw);
        }
      }
      if (k.use_z) {
        mw = warp_sum(mw);
        vw = warp_sum(vw);
        if (lane == 0) {
          k.mw[i] = mw;
          k.vw[i] = vw;
        }
      }
    }
  }
  double tot = block_sum_d(s, sh);
  if (tid == 0) {
    k.partS[blk] = tot;
    __threadfence();
    unsigned int ticket = atomicAdd(k.counter, 1u);
    last = (ticket == (unsigned)nblk - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    kl_finish(k, nblk, sh);
    if (tid == 0) *k.counter = 0u;
  }
}

struct KlPlan {
  KlDev k;
  dim3 grid;
};

static int kl_plan(mnf_ctx* ctx, const mnf_kl_args* a, float* kl_out, KlPlan* out) {
  MNF_CHECK_ARG(ctx && a && kl_out, "mnf_kl: NULL argument");
  MNF_CHECK_ARG(a->rows >= 1 && a->cols >= 1 && a->rows < (1 << 30) && a->cols < (1 << 30), "mnf_kl: bad shape");
  MNF_CHECK_ARG(a->mean_W && a->log_std_W && a->mean_b && a->log_var_b && a->prior_var_r_p && a->prior_var_r_p_bias,
                "mnf_kl: NULL parameter pointer");
  if (a->use_z) {
    MNF_CHECK_ARG(a->q0_log_var && a->r0_mean && a->r0_log_var && a->r0_apvar && a->z && a->ld_q && a->r_states &&
                      a->ld_r && a->n_r_states >= 1,
                  "mnf_kl: use_z needs q0/r0 parameters, z, ld_q, r_states, ld_r");
  }
  KlDev& k = out->k;
  memset(&k, 0, sizeof(k));
  k.conv = a->conv; k.use_z = a->use_z; k.r_all = a->r_all_states; k.n_states = a->n_r_states;
  k.rows = (int)a->rows; k.cols = (int)a->cols; k.D = a->conv ? k.cols : k.rows;
  k.Wm = a->mean_W; k.Ls = a->log_std_W; k.bm = a->mean_b; k.lvb = a->log_var_b;
  k.p = a->prior_var_r_p; k.pb = a->prior_var_r_p_bias;
  k.q0lv = a->q0_log_var; k.r0m = a->r0_mean; k.r0lv = a->r0_log_var; k.cap = a->r0_apvar;
  k.z = a->z; k.ldq = a->ld_q; k.rst = a->r_states; k.ldr = a->ld_r;
  k.eps_a = make_noise(&a->eps_a); k.eps_b = make_noise(&a->eps_b);
  k.kl_out = kl_out;
  if (a->conv) {
    out->grid = dim3((unsigned)cdiv(k.rows, KL_THREADS / 32), 1);
    k.nchunk = 0;
  } else {
    out->grid = dim3((unsigned)cdiv(k.rows, KL_ROWS), (unsigned)cdiv(k.cols, KL_THREADS));
    k.nchunk = (int)out->grid.x;
  }
  const int nblk = out->grid.x * out->grid.y;
  const int L = a->conv ? k.rows : k.cols;
  // workspace carve-up (8-byte aligned first)
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
  size_t o_partS = take(sizeof(double) * nblk);
  size_t o_pmw = take(sizeof(float) * (size_t)k.nchunk * k.cols);
  size_t o_pvw = take(sizeof(float) * (size_t)k.nchunk * k.cols);
  size_t o_mw = take(sizeof(float) * L);
  size_t o_vw = take(sizeof(float) * L);
  size_t o_aux = take(sizeof(float) * (L + 8));
  size_t o_dmw = take(sizeof(float) * L);
  size_t o_dvw = take(sizeof(float) * L);
  size_t o_cnt = take(sizeof(unsigned int) * 4);
  char* ws = (char*)mnf_workspace(ctx, off);
  if (!ws) return MNF_ERR_CUDA;
  k.partS = (double*)(ws + o_partS);
  k.part_mw = (float*)(ws + o_pmw);
  k.part_vw = (float*)(ws + o_pvw);
  k.mw = (float*)(ws + o_mw);
  k.vw = (float*)(ws + o_vw);
  k.aux = (float*)(ws + o_aux);
  k.dmw = (float*)(ws + o_dmw);
  k.dvw = (float*)(ws + o_dvw);
  k.counter = (unsigned int*)(ws + o_cnt);
  return MNF_OK;
}

extern "C" int mnf_kl_forward(mnf_ctx* ctx, const mnf_kl_args* a, float* kl_out) {
  KlPlan p;
  int rc = kl_plan(ctx, a, kl_out, &p);
  if (rc) return rc;
  MNF_CUDA(cudaMemsetAsync(p.k.counter, 0, sizeof(unsigned int) * 4, ctx->stream));
  kl_fwd_kernel<<<p.grid, KL_THREADS, 0, ctx->stream>>>(p.k);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

// ------------------------------------------------------------------------------ backward
struct KlGradDev {
  float *Wm, *Ls, *bm, *lvb, *p, *pb, *q0lv, *r0m, *r0lv, *cap, *dz, *drst;
};

// scalar tail of the backward pass: one CTA (App. A-4 / A-5 derivatives)
__global__ void __launch_bounds__(KL_THREADS) kl_bwd_scalar(KlDev k, float scale, KlGradDev g) {
  __shared__ double sh[32];
  const int tid = threadIdx.x;
  const int L = k.conv ? k.rows : k.cols;
  const double pb = k.pb[0], epb = exp(pb);
  // prior_var_r_p_bias (learn_p)
  if (g.pb) {
    double s = 0;
    for (int j = tid; j < k.cols; j += blockDim.x) {
      double b = k.bm[j];
      if (k.conv && k.use_z) b *= k.z[j];
      s += 1.0 - (exp((double)k.lvb[j]) + b * b) / epb;
    }
    s = block_sum_d(s, sh);
    if (tid == 0) g.pb[0] = (float)(scale * 0.5 * s);
  }
  if (!k.use_z) {
    for (int j = tid; j < k.cols; j += blockDim.x) {
      if (g.bm) g.bm[j] = (float)(scale * k.bm[j] / epb);
      if (g.lvb) g.lvb[j] = (float)(scale * 0.5 * (-1.0 + exp((double)k.lvb[j]) / epb));
    }
    return;
  }
  const double abar = k.aux[L], var_b = k.aux[L + 2];
  // Gaussian r-term
  double dab = 0;
  const int s0 = k.r_all ? 0 : k.n_states - 1;
  for (int d = tid; d < k.D; d += blockDim.x) {
    double mr = abar * k.r0m[d], lv = abar * k.r0lv[d], e = exp(lv);
    double s1 = 0, s2 = 0;
    for (int st = 0; st < k.n_states; ++st) {
      double diff = (double)k.rst[(size_t)st * k.D + d] - mr;
      bool in = st >= s0;
      if (in) { s1 += e * diff; s2 += -e * diff * diff + 1.0; }
      if (g.drst) g.drst[(size_t)st * k.D + d] = in ? (float)(scale * e * diff) : 0.f;
    }
    double dmr = -scale * s1, dlv = -scale * 0.5 * s2;
    if (g.r0m) g.r0m[d] = (float)(dmr * abar);
    if (g.r0lv) g.r0lv[d] = (float)(dlv * abar);
    if (g.q0lv) g.q0lv[d] = -0.5f * scale;
    dab += dmr * k.r0m[d] + dlv * k.r0lv[d];
  }
  const double dabar = block_sum_d(dab, sh);
  const double da = dabar / (double)L;
  for (int j = tid; j < L; j += blockDim.x) {
    float ea = noise_normal1(k.eps_a, 0, 0u, (uint32_t)j, 1, L);
    double dpre = da;
    if (!k.conv) {
      double a = k.aux[j];
      dpre = da * (1.0 - a * a);
    }
    k.dmw[j] = (float)dpre;
    k.dvw[j] = (float)(dpre * ea / (2.0 * sqrt((double)k.vw[j])));
  }
  if (!k.conv) {
    for (int j = tid; j < k.cols; j += blockDim.x) {
      if (g.bm) g.bm[j] = (float)(scale * k.bm[j] / epb);
      if (g.lvb) g.lvb[j] = (float)(scale * 0.5 * (-1.0 + exp((double)k.lvb[j]) / epb));
    }
  } else {
    float eb = noise_normal1(k.eps_b, 0, 0u, 0u, 1, 1);
    const double dmu_b = dabar, dvar_b = dabar * eb / (2.0 * sqrt(var_b));
    for (int f = tid; f < k.cols; f += blockDim.x) {
      double c = k.cap[f], zf = k.z[f], bmf = k.bm[f], bt = bmf * zf, elv = exp((double)k.lvb[f]);
      double dbt = scale * bt / epb + dmu_b * c;
      if (g.lvb) g.lvb[f] = (float)(scale * 0.5 * (-1.0 + elv / epb) + dvar_b * elv * c * c);
      if (g.bm) g.bm[f] = (float)(dbt * zf);
      if (g.dz) g.dz[f] = (float)(dbt * bmf);
      if (g.cap) g.cap[f] = (float)(dmu_b * bt + dvar_b * 2.0 * c * elv);
    }
  }
}

// element-wise pass over the weights: warp per row
__global__ void __launch_bounds__(KL_THREADS) kl_bwd_weights(KlDev k, float scale, KlGradDev g) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * (KL_THREADS / 32) + wid;
  if (i >= k.rows) return;
  const float pi = k.p[i], Pi = expf(pi);
  float rs_dz = 0.f, rs_dc = 0.f, rs_dp = 0.f;
  const float zi = (!k.conv && k.use_z) ? k.z[i] : 1.f;
  const float ci = (!k.conv && k.use_z) ? k.cap[i] : 0.f;
  const float dmw_i = (k.conv && k.use_z) ? k.dmw[i] : 0.f, dvw_i = (k.conv && k.use_z) ? k.dvw[i] : 0.f;
  for (int j = lane; j < k.cols; j += 32) {
    const size_t o = (size_t)i * k.cols + j;
    float ls = __ldg(k.Ls + o), wm = __ldg(k.Wm + o);
    float e = expf(ls), vt = e * e;
    float gw, gl;
    if (!k.conv) {
      float mt = zi * wm;
      float dmw = k.use_z ? k.dmw[j] : 0.f, dvw = k.use_z ? k.dvw[j] : 0.f;
      gw = scale * mt * zi / Pi + zi * ci * dmw;
      gl = scale * (-1.f + vt / Pi) + 2.f * vt * ci * ci * dvw;
      rs_dz += scale * mt * wm / Pi + wm * dmw * ci;
      rs_dc += mt * dmw + 2.f * ci * vt * dvw;
      rs_dp += 1.f - (vt + mt * mt) / Pi;
    } else {
      float zf = k.use_z ? k.z[j] : 1.f, cf = k.use_z ? k.cap[j] : 0.f;
      float mt = zf * wm;
      gw = scale * mt * zf / Pi + dmw_i * zf * cf;
      gl = scale * (-0.5f + vt / Pi) + 2.f * vt * dvw_i * cf * cf;
      rs_dp += 1.f - (vt + mt * mt) / Pi;
      if (k.use_z) {
        if (g.dz) atomicAdd(g.dz + j, scale * mt * wm / Pi + wm * dmw_i * cf);
        if (g.cap) atomicAdd(g.cap + j, mt * dmw_i + 2.f * cf * vt * dvw_i);
      }
    }
    if (g.Wm) g.Wm[o] = gw;
    if (g.Ls) g.Ls[o] = gl;
  }
  rs_dp = warp_sum(rs_dp);
  if (!k.conv) {
    rs_dz = warp_sum(rs_dz);
    rs_dc = warp_sum(rs_dc);
  }
  if (lane == 0) {
    if (g.p) g.p[i] = scale * 0.5f * rs_dp;
    if (!k.conv && k.use_z) {
      if (g.dz) g.dz[i] = rs_dz;
      if (g.cap) g.cap[i] = rs_dc;
    }
  }
}

extern "C" int mnf_kl_backward(mnf_ctx* ctx, const mnf_kl_args* a, float scale, mnf_kl_grads* gr) {
  MNF_CHECK_ARG(gr, "mnf_kl_backward: grads is NULL");
  KlPlan p;
  // the forward pass is re-run for its side outputs (mean_w, var_w, a, abar); its scalar
  // result lands in a scratch slot of the workspace
  float dummy_slot;
  (void)dummy_slot;
  int rc = kl_plan(ctx, a, (float*)1, &p);
  if (rc) return rc;
  p.k.kl_out = p.k.aux + ((a->conv ? a->rows : a->cols) + 4);
  MNF_CUDA(cudaMemsetAsync(p.k.counter, 0, sizeof(unsigned int) * 4, ctx->stream));
  kl_fwd_kernel<<<p.grid, KL_THREADS, 0, ctx->stream>>>(p.k);
  MNF_LAUNCHED(ctx);
  KlGradDev g = {gr->mean_W, gr->log_std_W, gr->mean_b, gr->log_var_b, gr->prior_var_r_p, gr->prior_var_r_p_bias,
                 gr->q0_log_var, gr->r0_mean, gr->r0_log_var, gr->r0_apvar, gr->d_z, gr->d_r_states};
  kl_bwd_scalar<<<1, KL_THREADS, 0, ctx->stream>>>(p.k, scale, g);
  MNF_LAUNCHED(ctx);
  kl_bwd_weights<<<(unsigned)cdiv(p.k.rows, KL_THREADS / 32), KL_THREADS, 0, ctx->stream>>>(p.k, scale, g);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}
