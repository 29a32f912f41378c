// IAF / RNVP flow steps with the GENERAL coupling networks the reference can build:
//   MADE   made.py:77-88   [MaskedDense(h_l), ReLU] for every entry of h_sizes, then MaskedDense(2 D)
//   RNVP   rnvp.py:20-31   Dense(h_l, activation) for every entry of h_sizes, then the t and s heads
// The fused kernels (flows_tile.cuh, flows_umma.cu, flow_chain_row1) cover what the MNF layers use --
// ONE hidden layer, ReLU / tanh -- and are the hot path.  Anything else (more hidden layers, another
// activation) runs here: one warp per row, the row's activations in shared memory, weights read through
// L1 / L2, forward (maf.py:47-53, rnvp.py:33-47) and backward with the same
// arithmetic contract (fp32, parity <= 1e-5 against the oracle).  Correct and reasonably parallel, not
// tuned: weight gradients are accumulated with atomics.
#include "common.cuh"

namespace fg {

constexpr int GW = 4;      // warps (rows) per CTA
constexpr int GH = 256;    // widest hidden layer
constexpr int GDMAX = 1024;

struct Args {
  int64_t B;
  int D, L, kind, parity, act, htot;
  int h[MNF_FLOW_MAX_HIDDEN];
  const float* w[MNF_FLOW_MAX_HIDDEN];
  const float* b[MNF_FLOW_MAX_HIDDEN];
  const float* m[MNF_FLOW_MAX_HIDDEN];
  const float *ws, *bs, *wt, *bt, *m2;
  int64_t ld2;
  const float* zin;
  float* zout;
  const float* ld_in;
  float* ld_out;
  float* hid;  // [B, htot] hidden activations (forward: written when non-null; backward: read)
  NoiseDev bern;
  // backward
  const float* dy;
  const float* dld;
  const float* add;
  float* dx;
  float* gw[MNF_FLOW_MAX_HIDDEN];
  float* gb[MNF_FLOW_MAX_HIDDEN];
  float *gws, *gbs, *gwt, *gbt;
};

__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case MNF_ACTV_RELU: return fmaxf(x, 0.f);
    case MNF_ACTV_TANH: return tanhf(x);
    case MNF_ACTV_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
// derivative in terms of the activation's OUTPUT h
__device__ __forceinline__ float act_bwd(int act, float h) {
  switch (act) {
    case MNF_ACTV_RELU: return h > 0.f ? 1.f : 0.f;
    case MNF_ACTV_TANH: return 1.f - h * h;
    case MNF_ACTV_SIGMOID: return h * (1.f - h);
    default: return 1.f;
  }
}

// shared memory of one warp: in[D] (masked input), mk[D] (RNVP mask), two D-wide scratch rows, two hidden buffers
struct Smem {
  float *in, *mk, *s1, *s2, *ha, *hb;
};
__device__ __forceinline__ Smem carve(float* base, int warp, int D) {
  float* p = base + (size_t)warp * (4 * (size_t)D + 2 * GH);
  Smem s;
  s.in = p; s.mk = p + D; s.s1 = p + 2 * D; s.s2 = p + 3 * D; s.ha = p + 4 * D; s.hb = s.ha + GH;
  return s;
}

// hidden stack: in[D] -> h_0 .. h_{L-1}; returns the buffer holding h_{L-1}.  `store`: hidden activations -> a.hid
__device__ __forceinline__ float* hidden_forward(const Args& a, const Smem& sm, int64_t row, int lane, bool store) {
  const float* src = sm.in;
  int n_in = a.D, off = 0;
  float* dst = sm.ha;
  for (int l = 0; l < a.L; ++l) {
    const int n_out = a.h[l];
    const float* W = a.w[l];
    const float* M = a.m[l];
    for (int j = lane; j < n_out; j += 32) {
      float acc = a.b[l][j];
      for (int i = 0; i < n_in; ++i) {
        float wv = W[(int64_t)i * n_out + j];
        if (M) wv *= M[(int64_t)i * n_out + j];
        acc = fmaf(src[i], wv, acc);
      }
      const float hv = act_fwd(a.act, acc);
      dst[j] = hv;
      if (store && a.hid) a.hid[row * a.htot + off + j] = hv;
    }
    __syncwarp();
    off += n_out;
    src = dst;
    n_in = n_out;
    dst = dst == sm.ha ? sm.hb : sm.ha;
  }
  return const_cast<float*>(src);
}

// masked input row (and the RNVP mask) -> shared memory
__device__ __forceinline__ void stage_input(const Args& a, const Smem& sm, int64_t row, int lane) {
  for (int i = lane; i < a.D; i += 32) {
    const float zv = a.zin[row * a.D + i];
    float mk = 1.f;
    if (a.kind == MNF_FLOW_RNVP) mk = noise_bern1(a.bern, row, (uint32_t)i, a.D, 0.5f);
    sm.mk[i] = mk;
    sm.in[i] = a.kind == MNF_FLOW_RNVP ? mk * zv : zv;
  }
  __syncwarp();
}

// (s, t) of output column d from the last hidden layer
__device__ __forceinline__ void heads(const Args& a, const float* hl, int d, float& s, float& t) {
  const int H = a.h[a.L - 1];
  s = a.bs[d];
  t = a.bt[d];
  for (int j = 0; j < H; ++j) {
    float ws = a.ws[(int64_t)j * a.ld2 + d], wt = a.wt[(int64_t)j * a.ld2 + d];
    if (a.m2) {  // IAF: one [H, 2D] mask, s columns then t columns
      ws *= a.m2[(int64_t)j * a.ld2 + d];
      wt *= a.m2[(int64_t)j * a.ld2 + a.D + d];
    }
    s = fmaf(hl[j], ws, s);
    t = fmaf(hl[j], wt, t);
  }
}

__global__ void __launch_bounds__(GW * 32) generic_fwd(Args a) {
  extern __shared__ float gsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Smem sm = carve(gsm, warp, a.D);
  for (int64_t row = (int64_t)blockIdx.x * GW + warp; row < a.B; row += (int64_t)gridDim.x * GW) {
    stage_input(a, sm, row, lane);
    const float* hl = hidden_forward(a, sm, row, lane, true);
    float ld = 0.f;
    for (int d = lane; d < a.D; d += 32) {
      float s, t;
      heads(a, hl, d, s, t);
      const float zv = a.zin[row * a.D + d];
      if (a.kind == MNF_FLOW_IAF) {
        const float y = fmaf(zv, expf(s), t);  // maf.py:50
        a.zout[row * a.D + (a.parity ? a.D - 1 - d : d)] = y;
        ld += s;
      } else {
        const float m = sm.mk[d];
        const float g = 1.f / (1.f + expf(-s));  // rnvp.py:41-46
        a.zout[row * a.D + d] = ((1.f - m) * zv * g + (1.f - g) * t) + m * zv;
        ld += (1.f - m) * logf(g);
      }
    }
    ld = warp_sum(ld);
    if (lane == 0 && a.ld_out) a.ld_out[row] = (a.ld_in ? a.ld_in[row] : 0.f) + ld;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(GW * 32) generic_bwd(Args a) {
  extern __shared__ float gsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Smem sm = carve(gsm, warp, a.D);
  const int D = a.D, HL = a.h[a.L - 1];
  for (int64_t row = (int64_t)blockIdx.x * GW + warp; row < a.B; row += (int64_t)gridDim.x * GW) {
    stage_input(a, sm, row, lane);
    // the saved hidden activations of the last layer (everything else is read from a.hid as needed)
    const float* hid = a.hid + row * a.htot;
    const float* hl = hid + (a.htot - HL);
    const float dldr = a.dld ? a.dld[row] : 0.f;
    // ---- heads: ds (s1), dt (s2); direct dx parked in a.dx
    for (int d = lane; d < D; d += 32) {
      float s, t;
      heads(a, hl, d, s, t);
      const float zv = a.zin[row * D + d];
      float ds, dt, dxd;
      if (a.kind == MNF_FLOW_IAF) {
        const float g = a.dy[row * D + (a.parity ? D - 1 - d : d)];
        const float es = expf(s);
        ds = g * zv * es + dldr;
        dt = g;
        dxd = g * es;
      } else {
        const float g_up = a.dy[row * D + d], m = sm.mk[d];
        const float gate = 1.f / (1.f + expf(-s));
        const float dg = g_up * ((1.f - m) * zv - t) + dldr * (1.f - m) / gate;
        ds = dg * gate * (1.f - gate);
        dt = g_up * (1.f - gate);
        dxd = g_up * ((1.f - m) * gate + m);
      }
      sm.s1[d] = ds;
      sm.s2[d] = dt;
      a.dx[row * D + d] = dxd + (a.add ? a.add[row * D + d] : 0.f);
      atomicAdd(a.gbs + d, ds);
      atomicAdd(a.gbt + d, dt);
      for (int j = 0; j < HL; ++j) {  // head weight gradients, coalesced over d
        float gs = hl[j] * ds, gt = hl[j] * dt;
        if (a.m2) {
          gs *= a.m2[(int64_t)j * a.ld2 + d];
          gt *= a.m2[(int64_t)j * a.ld2 + D + d];
        }
        atomicAdd(a.gws + (int64_t)j * a.ld2 + d, gs);
        atomicAdd(a.gwt + (int64_t)j * a.ld2 + d, gt);
      }
    }
    __syncwarp();
    // ---- dh of the last hidden layer
    float* dh = sm.ha;
    float* dh2 = sm.hb;
    for (int j = lane; j < HL; j += 32) {
      float acc = 0.f;
      for (int d = 0; d < D; ++d) {
        float ws = a.ws[(int64_t)j * a.ld2 + d], wt = a.wt[(int64_t)j * a.ld2 + d];
        if (a.m2) {
          ws *= a.m2[(int64_t)j * a.ld2 + d];
          wt *= a.m2[(int64_t)j * a.ld2 + D + d];
        }
        acc = fmaf(sm.s1[d], ws, fmaf(sm.s2[d], wt, acc));
      }
      dh[j] = acc;
    }
    __syncwarp();
    // ---- hidden stack, last layer first
    int off = a.htot;
    for (int l = a.L - 1; l >= 0; --l) {
      const int n_out = a.h[l], n_in = l > 0 ? a.h[l - 1] : D;
      off -= n_out;
      const float* hcur = hid + off;
      const float* hin = l > 0 ? hid + off - n_in : sm.in;
      const float* W = a.w[l];
      const float* M = a.m[l];
      for (int j = lane; j < n_out; j += 32) {  // dpre in place of dh; bias and kernel gradients (coalesced over j)
        const float dp = dh[j] * act_bwd(a.act, hcur[j]);
        dh[j] = dp;
        atomicAdd(a.gb[l] + j, dp);
        for (int i = 0; i < n_in; ++i) {
          float g = hin[i] * dp;
          if (M) g *= M[(int64_t)i * n_out + j];
          atomicAdd(a.gw[l] + (int64_t)i * n_out + j, g);
        }
      }
      __syncwarp();
      float* dprev = l > 0 ? dh2 : sm.s1;  // d(input of layer l): next dh, or the input-row gradient
      for (int i = lane; i < n_in; i += 32) {
        float acc = 0.f;
        for (int j = 0; j < n_out; ++j) {
          float wv = W[(int64_t)i * n_out + j];
          if (M) wv *= M[(int64_t)i * n_out + j];
          acc = fmaf(dh[j], wv, acc);
        }
        dprev[i] = acc;
      }
      __syncwarp();
      if (l > 0) {
        float* t = dh; dh = dh2; dh2 = t;
      }
    }
    // ---- through the (masked) input
    for (int d = lane; d < D; d += 32) a.dx[row * D + d] += (a.kind == MNF_FLOW_RNVP ? sm.mk[d] : 1.f) * sm.s1[d];
    __syncwarp();
  }
}


// MAF.forward = IAF.inverse, maf.py:33-45: the sequential decode.  As written: x = 0; z reversed when
// parity; for i in range(D): (s, t) = net(x); x[:, i] = (z[:, i] - t[:, i]) exp(-s[:, i]); log_det -= s[:, i].
// (The reference assigns into a tensor slice, which TensorFlow rejects -- App. B-4; this is the loop it
// spells out.)  One warp per row.  net(x) changes by one input per iteration, so the first layer's
// pre-activation is UPDATED (p0 += x_i W0[i, :]) instead of recomputed: O(D (h0 + sum h_l h_l+1)) per row
// instead of D full passes; the remaining layers and the two head columns i are evaluated per iteration.
__global__ void __launch_bounds__(GW * 32) maf_decode_kernel(Args a) {
  extern __shared__ float gsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per warp: p0[GH] running first-layer pre-activation, two hidden buffers
  float* p0 = gsm + (size_t)warp * 3 * GH;
  float* ha = p0 + GH;
  float* hb = ha + GH;
  const int D = a.D, H0 = a.h[0], HL = a.h[a.L - 1];
  for (int64_t row = (int64_t)blockIdx.x * GW + warp; row < a.B; row += (int64_t)gridDim.x * GW) {
    for (int j = lane; j < H0; j += 32) p0[j] = a.b[0][j];
    __syncwarp();
    float ld = 0.f;
    for (int i = 0; i < D; ++i) {
      // hidden stack from the current pre-activation
      for (int j = lane; j < H0; j += 32) ha[j] = act_fwd(a.act, p0[j]);
      __syncwarp();
      float* src = ha;
      float* dst = hb;
      int n_in = H0;
      for (int l = 1; l < a.L; ++l) {
        const int n_out = a.h[l];
        for (int j = lane; j < n_out; j += 32) {
          float acc = a.b[l][j];
          for (int q = 0; q < n_in; ++q) {
            float wv = a.w[l][(int64_t)q * n_out + j];
            if (a.m[l]) wv *= a.m[l][(int64_t)q * n_out + j];
            acc = fmaf(src[q], wv, acc);
          }
          dst[j] = act_fwd(a.act, acc);
        }
        __syncwarp();
        float* t = src; src = dst; dst = t;
        n_in = n_out;
      }
      // head columns i: lanes split the last hidden layer
      float ps = 0.f, pt = 0.f;
      for (int j = lane; j < HL; j += 32) {
        float ws = a.ws[(int64_t)j * a.ld2 + i], wt = a.wt[(int64_t)j * a.ld2 + i];
        if (a.m2) {
          ws *= a.m2[(int64_t)j * a.ld2 + i];
          wt *= a.m2[(int64_t)j * a.ld2 + D + i];
        }
        ps = fmaf(src[j], ws, ps);
        pt = fmaf(src[j], wt, pt);
      }
      const float s = warp_sum(ps) + a.bs[i], t = warp_sum(pt) + a.bt[i];
      const float zi = a.zin[row * D + (a.parity ? D - 1 - i : i)];
      const float xi = (zi - t) * expf(-s);
      ld -= s;
      if (lane == 0) a.zout[row * D + i] = xi;
      for (int j = lane; j < H0; j += 32) {
        float wv = a.w[0][(int64_t)i * H0 + j];
        if (a.m[0]) wv *= a.m[0][(int64_t)i * H0 + j];
        p0[j] = fmaf(xi, wv, p0[j]);
      }
      __syncwarp();
    }
    if (lane == 0 && a.ld_out) a.ld_out[row] = ld;
  }
}

int fill(const mnf_flow_step& st, int64_t B, int64_t D, Args& a) {
  MNF_CHECK_ARG(D <= GDMAX, "general flow networks (several hidden layers / other activations) support D <= %d, got %lld", GDMAX,
                (long long)D);
  MNF_CHECK_ARG(st.n_extra >= 0 && st.n_extra < MNF_FLOW_MAX_HIDDEN, "flow step: n_extra=%d not in [0,%d)", st.n_extra,
                MNF_FLOW_MAX_HIDDEN);
  MNF_CHECK_ARG(st.activation >= MNF_ACTV_DEFAULT && st.activation <= MNF_ACTV_LINEAR, "flow step: unknown activation %d", st.activation);
  MNF_CHECK_ARG(st.kind != MNF_FLOW_IAF || st.activation == MNF_ACTV_DEFAULT || st.activation == MNF_ACTV_RELU,
                "MADE hidden layers are ReLU (made.py:77-88): activation %d", st.activation);
  a = Args{};
  a.B = B; a.D = (int)D; a.L = 1 + st.n_extra; a.kind = st.kind; a.parity = st.parity;
  a.act = st.activation != MNF_ACTV_DEFAULT ? st.activation : (st.kind == MNF_FLOW_IAF ? MNF_ACTV_RELU : MNF_ACTV_TANH);
  a.h[0] = st.hidden; a.w[0] = st.w1; a.b[0] = st.b1; a.m[0] = st.m1;
  for (int l = 0; l < st.n_extra; ++l) {
    a.h[l + 1] = st.extra_h[l]; a.w[l + 1] = st.extra_w[l]; a.b[l + 1] = st.extra_b[l]; a.m[l + 1] = st.extra_m[l];
  }
  a.htot = 0;
  for (int l = 0; l < a.L; ++l) {
    MNF_CHECK_ARG(a.h[l] >= 1 && a.h[l] <= GH && a.w[l] && a.b[l], "flow step: hidden layer %d (width %d, max %d) incomplete", l, a.h[l], GH);
    a.htot += a.h[l];
  }
  MNF_CHECK_ARG(st.ws && st.bs && st.wt && st.bt, "flow step: NULL head weights");
  a.ws = st.ws; a.bs = st.bs; a.wt = st.wt; a.bt = st.bt; a.m2 = st.m2; a.ld2 = st.ld2;
  a.bern = make_noise(&st.bern);
  return MNF_OK;
}

size_t smem_bytes(int D) { return sizeof(float) * GW * (4 * (size_t)D + 2 * GH); }

template <typename K>
int set_smem(K kern, size_t bytes) {
  MNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return MNF_OK;
}

}  // namespace fg

bool mnf_flow_step_is_generic(const mnf_flow_step& st) {
  if (st.kind != MNF_FLOW_IAF && st.kind != MNF_FLOW_RNVP) return false;
  if (st.n_extra > 0) return true;
  if (st.activation == MNF_ACTV_DEFAULT) return false;
  return !((st.kind == MNF_FLOW_IAF && st.activation == MNF_ACTV_RELU) || (st.kind == MNF_FLOW_RNVP && st.activation == MNF_ACTV_TANH));
}

int64_t mnf_flow_step_hidden_total(const mnf_flow_step& st) {
  if (st.kind == MNF_FLOW_PLANAR) return 0;
  int64_t h = st.hidden;
  for (int l = 0; l < st.n_extra && l < MNF_FLOW_MAX_HIDDEN - 1; ++l) h += st.extra_h[l];
  return h;
}

int mnf_flow_generic_forward(mnf_ctx* ctx, int64_t B, int64_t D, const mnf_flow_step& st, const float* zin, float* zout,
                             const float* ld_in, float* ld_out, float* hid_save) {
  fg::Args a;
  int rc = fg::fill(st, B, D, a);
  if (rc) return rc;
  a.zin = zin; a.zout = zout; a.ld_in = ld_in; a.ld_out = ld_out; a.hid = hid_save;
  const size_t smem = fg::smem_bytes((int)D);
  rc = fg::set_smem(fg::generic_fwd, smem);
  if (rc) return rc;
  const int64_t blocks = cdiv(B, fg::GW), cap = (int64_t)ctx->sm_count * 8;
  mnf_note_path("flow_generic");
  fg::generic_fwd<<<(unsigned)(blocks < cap ? blocks : cap), fg::GW * 32, smem, ctx->stream>>>(a);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

// dx = d(step)/d(input) applied to dy (+ add); parameter gradients are ACCUMULATED into gs's slots
int mnf_flow_generic_backward(mnf_ctx* ctx, int64_t B, int64_t D, const mnf_flow_step& st, const mnf_flow_step& gs,
                              const float* xin, const float* hid, const float* dy, const float* dld, const float* add,
                              float* dx) {
  fg::Args a;
  int rc = fg::fill(st, B, D, a);
  if (rc) return rc;
  MNF_CHECK_ARG(hid, "mnf_flow_backward: saved hidden activations required");
  MNF_CHECK_ARG(gs.w1 && gs.b1 && gs.ws && gs.bs && gs.wt && gs.bt, "flow step: NULL grad pointer");
  a.zin = xin; a.hid = const_cast<float*>(hid); a.dy = dy; a.dld = dld; a.add = add; a.dx = dx;
  a.gw[0] = (float*)gs.w1; a.gb[0] = (float*)gs.b1;
  for (int l = 0; l < st.n_extra; ++l) {
    MNF_CHECK_ARG(gs.extra_w[l] && gs.extra_b[l], "flow step: NULL grad pointer (hidden layer %d)", l + 1);
    a.gw[l + 1] = (float*)gs.extra_w[l]; a.gb[l + 1] = (float*)gs.extra_b[l];
  }
  a.gws = (float*)gs.ws; a.gbs = (float*)gs.bs; a.gwt = (float*)gs.wt; a.gbt = (float*)gs.bt;
  const size_t smem = fg::smem_bytes((int)D);
  rc = fg::set_smem(fg::generic_bwd, smem);
  if (rc) return rc;
  const int64_t blocks = cdiv(B, fg::GW), cap = (int64_t)ctx->sm_count * 8;
  fg::generic_bwd<<<(unsigned)(blocks < cap ? blocks : cap), fg::GW * 32, smem, ctx->stream>>>(a);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

extern "C" int mnf_maf_decode(mnf_ctx* ctx, int64_t B, int64_t D, const mnf_flow_step* step, const float* z, float* x,
                              float* log_dets) {
  MNF_CHECK_ARG(ctx && step && z && x, "mnf_maf_decode: NULL argument");
  MNF_CHECK_ARG(step->kind == MNF_FLOW_IAF, "mnf_maf_decode: a MADE-based (MAF / IAF) step is required");
  MNF_CHECK_ARG(B >= 0 && D >= 1, "mnf_maf_decode: bad shape");
  if (B == 0) return MNF_OK;
  fg::Args a;
  int rc = fg::fill(*step, B, D, a);
  if (rc) return rc;
  a.zin = z; a.zout = x; a.ld_out = log_dets;
  const size_t smem = sizeof(float) * fg::GW * 3 * fg::GH;
  const int64_t blocks = cdiv(B, fg::GW), cap = (int64_t)ctx->sm_count * 8;
  mnf_note_path("maf_decode");
  fg::maf_decode_kernel<<<(unsigned)(blocks < cap ? blocks : cap), fg::GW * 32, smem, ctx->stream>>>(a);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}
