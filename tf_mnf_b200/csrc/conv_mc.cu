// Monte-Carlo predictive through a FIRST MNFConv2D layer ("rows sharing an image
// share conv(x, W)").  MNFConv2D.call (mnf_conv.py:182-197) applies the multiplicative noise AFTER
// the contractions:  y = (conv(x, W_mu) + b_mu) * z + sqrt(conv(x^2, sigma^2) + vb) * eps,  so for a
// batch that repeats every image n times (lenet_mnist.py:84, x.repeat(...)) the two convolutions
// are the same for all n posterior samples of an image.  conv_mc_maps computes the pre-noise
// mean and standard-deviation maps once per DISTINCT image; conv_mc_expand then draws
// z-scaled, eps-perturbed outputs for every (image, sample) row -- the part that is really
// per-row -- optionally with the ReLU + 2x2 max-pool that follows the layer in MNF-LeNet.
// Noise counters are those of the full-batch kernels (global row, output pixel, filter quad), so
// the draws are the same numbers either way.
//
// One CTA per output row: the per-row constants (image, z, Philox row) are set up once and the
// threads walk the row's (pooled pixel, filter quad) items; the per-image maps (460 KB for
// MNF-LeNet conv1) stay in L2 / L1.  Issue-bound on Philox4x32-10 + Box-Muller like every noise
// consumer here, but with every warp of the SM on it instead of the 16 epilogue warps of a GEMM CTA.
#include "common.cuh"

namespace {

struct McArgs {
  int64_t B, n_rep;
  int Ho, Wo, F;
  const float* mean;  // [I, Ho, Wo, F]  conv(x, W_mu) + b_mu
  const float* sd;    // [I, Ho, Wo, F]  sqrt(conv(x^2, sigma^2) + vb)
  const float* z;     // [B, F] or nullptr
  NoiseDev eps;
  float* y;           // [B, Ho, Wo, F] or pooled [B, Ho/2, Wo/2, F]
};

__device__ __forceinline__ float4 ld4g(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// four outputs (one pixel, one filter quad) of a row; `off` = pix * F + 4 q indexes the row's injected
// noise and the image's maps alike (32-bit: the per-quad 64-bit address arithmetic was 10 % of the kernel)
template <bool INJ>
__device__ __forceinline__ void mc_quad(const McArgs& a, const PhiloxKeys& K, const float* inj, uint64_t grow, const float* mimg,
                                        const float* simg, const float4& zq, int pix, int q, int off, float o[4]) {
  float e[4];
  if (INJ) {
    const float4 t = ld4g(inj + off);
    e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w;
  } else {
    philox_normal4_keys(a.eps, K, grow, (uint32_t)pix, (uint32_t)q, e);
  }
  const float4 m = ld4g(mimg + off), s = ld4g(simg + off);
  o[0] = fmaf(s.x, e[0], m.x * zq.x);
  o[1] = fmaf(s.y, e[1], m.y * zq.y);
  o[2] = fmaf(s.z, e[2], m.z * zq.z);
  o[3] = fmaf(s.w, e[3], m.w * zq.w);
}

// Pre-noise maps of the DISTINCT images: mean = conv(x, W_mu) + b_mu, sd = sqrt(conv(x^2, s^2) + vb)
// with s = min(exp(L_sigma), max_std), vb = min(exp(l_vb), max_std^2) (mnf_conv.py:184-191).  A handful
// of images (10 in the MC-predictive config): one thread per output element walks the K = kh*kw*C
// filter taps; the general kernels would run it on one or two CTAs (116 us on the critical path).
struct MapArgs {
  int64_t I;
  int H, W, C, kh, kw, F, Ho, Wo, stride, pad_t, pad_l;
  const float *x, *Wm, *Ls, *bm, *lvb;
  float max_std;
  float *mean, *sd;
};
__global__ void __launch_bounds__(128) conv_mc_maps(MapArgs a) {
  const int64_t total = a.I * a.Ho * a.Wo * a.F;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(idx % a.F);
    int64_t t = idx / a.F;
    const int wo = (int)(t % a.Wo);
    t /= a.Wo;
    const int ho = (int)(t % a.Ho);
    const int64_t img = t / a.Ho;
    const float* xi = a.x + img * (int64_t)a.H * a.W * a.C;
    float m = 0.f, v = 0.f;
    for (int i = 0; i < a.kh; ++i) {
      const int hi = ho * a.stride - a.pad_t + i;
      if (hi < 0 || hi >= a.H) continue;
      for (int j = 0; j < a.kw; ++j) {
        const int wi = wo * a.stride - a.pad_l + j;
        if (wi < 0 || wi >= a.W) continue;
        const float* xp = xi + ((int64_t)hi * a.W + wi) * a.C;
        const int64_t wbase = ((int64_t)(i * a.kw + j) * a.C) * a.F + f;
        for (int c = 0; c < a.C; ++c) {
          const float xv = __ldg(xp + c);
          const float sg = fminf(expf(__ldg(a.Ls + wbase + (int64_t)c * a.F)), a.max_std);
          m = fmaf(xv, __ldg(a.Wm + wbase + (int64_t)c * a.F), m);
          v = fmaf(xv * xv, sg * sg, v);
        }
      }
    }
    a.mean[idx] = m + __ldg(a.bm + f);
    a.sd[idx] = sqrtf(v + fminf(expf(__ldg(a.lvb + f)), a.max_std * a.max_std));
  }
}

// INJ: the noise is read from memory (tests); compiled apart so that the Philox path is branch-free and the four
// independent Philox chains of a pooled item can be interleaved by the scheduler
template <bool POOL, bool INJ>
__global__ void __launch_bounds__(256) conv_mc_expand(McArgs a) {
  const int64_t brow = blockIdx.x;
  const int64_t img = brow / a.n_rep;
  const uint64_t grow = a.eps.row_offset + (uint64_t)brow;
  const int Q = a.F >> 2;
  const int64_t map = (int64_t)a.Ho * a.Wo * a.F;
  const float* mimg = a.mean + img * map;
  const float* simg = a.sd + img * map;
  const float* zrow = a.z ? a.z + brow * a.F : nullptr;
  const PhiloxKeys K = philox_keys(a.eps);
  const float* inj = a.eps.injected ? a.eps.injected + brow * map : nullptr;
  if (POOL) {
    const int Hp = a.Ho >> 1, Wp = a.Wo >> 1, items = Hp * Wp * Q;
    float* yrow = a.y + brow * (int64_t)items * 4;
    // (q, wp, hp) of item `it`, advanced incrementally (runtime integer divisions cost ~20 instructions each)
    const int sq = (int)blockDim.x % Q, spp = (int)blockDim.x / Q, swp = spp % Wp, shp = spp / Wp;
    int q = (int)threadIdx.x % Q, pp0 = (int)threadIdx.x / Q, wp = pp0 % Wp, hp = pp0 / Wp;
    const int dpix[4] = {0, 1, a.Wo, a.Wo + 1};
    const int doff[4] = {0, a.F, a.Wo * a.F, (a.Wo + 1) * a.F};
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const float4 zq = zrow ? ld4g(zrow + 4 * q) : make_float4(1.f, 1.f, 1.f, 1.f);
      float best[4] = {0.f, 0.f, 0.f, 0.f};  // max(., 0) folds the ReLU in
      const int pix0 = 2 * hp * a.Wo + 2 * wp, off0 = pix0 * a.F + 4 * q;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        float o[4];
        mc_quad<INJ>(a, K, inj, grow, mimg, simg, zq, pix0 + dpix[d], q, off0 + doff[d], o);
#pragma unroll
        for (int j = 0; j < 4; ++j) best[j] = fmaxf(best[j], o[j]);
      }
      *reinterpret_cast<float4*>(yrow + it * 4) = make_float4(best[0], best[1], best[2], best[3]);
      q += sq; wp += swp; hp += shp;
      if (q >= Q) { q -= Q; ++wp; }
      if (wp >= Wp) { wp -= Wp; ++hp; }
    }
  } else {
    const int items = a.Ho * a.Wo * Q;
    float* yrow = a.y + brow * (int64_t)items * 4;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int q = it % Q, pix = it / Q;
      const float4 zq = zrow ? ld4g(zrow + 4 * q) : make_float4(1.f, 1.f, 1.f, 1.f);
      float o[4];
      mc_quad<INJ>(a, K, inj, grow, mimg, simg, zq, pix, q, it * 4, o);
      *reinterpret_cast<float4*>(yrow + it * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace

extern "C" {

int mnf_conv2d_mc_can_expand(mnf_ctx* ctx, int64_t Ho, int64_t Wo, int64_t F, int32_t fuse_relu_pool, int32_t* ok) {
  MNF_CHECK_ARG(ctx && ok && Ho >= 1 && Wo >= 1 && F >= 1, "mnf_conv2d_mc_can_expand: bad argument");
  *ok = (F % 4 == 0 && Ho * Wo * F < (1 << 28) && (!fuse_relu_pool || (Ho % 2 == 0 && Wo % 2 == 0))) ? 1 : 0;
  return MNF_OK;
}

int mnf_conv2d_mc_maps(mnf_ctx* ctx, const mnf_conv_shape* s, const float* x, const float* mean_W, const float* log_std_W,
                       const float* mean_b, const float* log_var_b, float max_std, float* mean, float* sd) {
  MNF_CHECK_ARG(ctx && s && x && mean_W && log_std_W && mean_b && log_var_b && mean && sd, "mnf_conv2d_mc_maps: NULL argument");
  int64_t Ho, Wo;
  int rc = mnf_conv_out_hw(s, &Ho, &Wo);
  if (rc) return rc;
  if (s->B == 0) return MNF_OK;
  MapArgs a;
  a.I = s->B; a.H = (int)s->H; a.W = (int)s->W; a.C = (int)s->C; a.kh = (int)s->kh; a.kw = (int)s->kw; a.F = (int)s->F;
  a.Ho = (int)Ho; a.Wo = (int)Wo; a.stride = (int)s->stride; a.pad_t = a.pad_l = 0;
  if (s->same) {  // TF SAME: smaller half of the padding first
    int64_t ph = (Ho - 1) * s->stride + s->kh - s->H, pw = (Wo - 1) * s->stride + s->kw - s->W;
    a.pad_t = (int)((ph > 0 ? ph : 0) / 2);
    a.pad_l = (int)((pw > 0 ? pw : 0) / 2);
  }
  a.x = x; a.Wm = mean_W; a.Ls = log_std_W; a.bm = mean_b; a.lvb = log_var_b; a.max_std = max_std; a.mean = mean; a.sd = sd;
  const int64_t total = a.I * Ho * Wo * a.F;
  const int64_t cap = (int64_t)ctx->sm_count * 16;
  const int grid = (int)(cdiv(total, 128) < cap ? cdiv(total, 128) : cap);
  conv_mc_maps<<<grid, 128, 0, ctx->stream>>>(a);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

int mnf_conv2d_mc_expand(mnf_ctx* ctx, int64_t B, int64_t n_rep, int64_t Ho, int64_t Wo, int64_t F, const float* mean,
                         const float* sd, const float* z, const mnf_noise* eps, int32_t fuse_relu_pool, float* y) {
  MNF_CHECK_ARG(ctx && mean && sd && eps && y, "mnf_conv2d_mc_expand: NULL argument");
  MNF_CHECK_ARG(B >= 0 && n_rep >= 1 && Ho >= 1 && Wo >= 1 && F >= 1 && B < (1ll << 31), "mnf_conv2d_mc_expand: bad shape");
  int32_t ok = 0;
  int rc = mnf_conv2d_mc_can_expand(ctx, Ho, Wo, F, fuse_relu_pool, &ok);
  if (rc) return rc;
  if (!ok) {
    mnf_set_error("mnf_conv2d_mc_expand: unsupported shape Ho=%lld Wo=%lld F=%lld (F %% 4 == 0; even Ho, Wo when pooling)",
                  (long long)Ho, (long long)Wo, (long long)F);
    return MNF_ERR_UNSUPPORTED;
  }
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  MNF_CHECK_ARG(al(mean) && al(sd) && al(z) && al(y) && al(eps->injected), "mnf_conv2d_mc_expand: buffers must be 16-byte aligned");
  if (B == 0) return MNF_OK;
  McArgs a;
  a.B = B; a.n_rep = n_rep; a.Ho = (int)Ho; a.Wo = (int)Wo; a.F = (int)F;
  a.mean = mean; a.sd = sd; a.z = z; a.eps = make_noise(eps); a.y = y;
  mnf_note_path(fuse_relu_pool ? "conv_mc_expand<pool>" : "conv_mc_expand");
  const bool probed = mnf_probe_begin(ctx, MNF_PROBE_MC_EXPAND);
  const bool inj = a.eps.injected != nullptr;
  if (fuse_relu_pool) {
    if (inj) conv_mc_expand<true, true><<<(unsigned)B, 256, 0, ctx->stream>>>(a);
    else conv_mc_expand<true, false><<<(unsigned)B, 256, 0, ctx->stream>>>(a);
  } else {
    if (inj) conv_mc_expand<false, true><<<(unsigned)B, 256, 0, ctx->stream>>>(a);
    else conv_mc_expand<false, false><<<(unsigned)B, 256, 0, ctx->stream>>>(a);
  }
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

}  // extern "C"
