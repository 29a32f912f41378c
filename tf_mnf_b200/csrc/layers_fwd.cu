// MNFDense.call (mnf_dense.py:159-170) and MNFConv2D.call (mnf_conv.py:182-197) forward:
// the paired mean/variance contraction with the local-reparameterisation epilogue fused in.
//
// fp32-exact path (<= 1e-5 rel. vs the fp64 oracle): register-blocked FFMA GEMM / implicit
// GEMM.  One staged x tile feeds BOTH products: the loader writes a1 = x*z (dense) or x
// (conv) and a2 = x*x into shared memory; the CTA accumulates mean and variance tiles side
// by side and the epilogue adds the biases, applies the conv z-scale, draws eps from Philox
// (or reads the injected draw) and writes y once.
#include <cstdlib>
#include "common.cuh"

#define PG_BM 128
#define PG_BK 16
#define PG_TM 8
#define PG_TN 4
#define PG_PAD 4

struct PairedArgs {
  int64_t M;  // rows of the (implicit) GEMM: B (dense) or B*Ho*Wo (conv)
  int K, N;
  const float* x;
  const float* z;   // dense: [M,K]; conv: [B,N]; nullptr = ones
  const float* Wm;  // [K,N]
  const float* Vw;  // [K,N] clipped variance (precomputed)
  const float* bm;  // [N]
  const float* vb;  // [N] clipped bias variance
  float* y;
  float* mean_out;  // conv: pre-z mean (for dz in backward); may be nullptr
  float* sd_out;    // sqrt(V); may be nullptr
  NoiseDev eps;
  // conv geometry
  int H, W, C, kh, kw, Ho, Wo, stride, pad_t, pad_l;
  int pool;  // conv only: fuse ReLU + 2x2/stride-2 max-pool into the epilogue (y is [B,Ho/2,Wo/2,F])
};

// Vw = min(exp(Ls), max_std)^2 ; vb = min(exp(lvb), max_std^2)   (mnf_dense.py:163-165)
__global__ void prep_var_kernel(int64_t nW, const float* __restrict__ Ls, float max_std, float* __restrict__ Vw,
                                int N, const float* __restrict__ lvb, float* __restrict__ vb) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nW; i += (int64_t)gridDim.x * blockDim.x) {
    float s = fminf(expf(Ls[i]), max_std);
    Vw[i] = s * s;
    if (i < N) vb[i] = fminf(expf(lvb[i]), max_std * max_std);
  }
}

template <bool CONV, int BN>
__global__ void __launch_bounds__((PG_BM / PG_TM) * (BN / PG_TN)) paired_gemm_fwd(PairedArgs a) {
  constexpr int NT = (PG_BM / PG_TM) * (BN / PG_TN);
  constexpr int TXN = BN / PG_TN;
  constexpr int A_PER = PG_BM * PG_BK / NT;  // A elements per thread per k-step
  constexpr int B_PER = PG_BK * BN / NT;
  constexpr int AROWS = NT / PG_BK;          // rows covered per pass of the A loader

  __shared__ __align__(16) float As1[PG_BK][PG_BM + PG_PAD];
  __shared__ __align__(16) float As2[PG_BK][PG_BM + PG_PAD];
  __shared__ __align__(16) float Bs1[PG_BK][BN + PG_PAD];
  __shared__ __align__(16) float Bs2[PG_BK][BN + PG_PAD];
  __shared__ int64_t rowoff[CONV ? PG_BM : 1];
  __shared__ short rhi[CONV ? PG_BM : 1], rwi[CONV ? PG_BM : 1];

  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int64_t m0 = (int64_t)blockIdx.x * PG_BM;
  const int n0 = blockIdx.y * BN;
  const int K = a.K, N = a.N;

  if (CONV) {
    for (int m = tid; m < PG_BM; m += NT) {
      int64_t gm = m0 + m;
      if (gm < a.M) {
        int pix = (int)(gm % ((int64_t)a.Ho * a.Wo));
        int64_t b = gm / ((int64_t)a.Ho * a.Wo);
        int ho = pix / a.Wo, wo = pix - ho * a.Wo;
        int hi0 = ho * a.stride - a.pad_t, wi0 = wo * a.stride - a.pad_l;
        rowoff[m] = ((b * a.H + hi0) * a.W + wi0) * a.C;
        rhi[m] = (short)hi0;
        rwi[m] = (short)wi0;
      } else {
        rowoff[m] = 0;
        rhi[m] = (short)-20000;  // forces the bounds check to fail
        rwi[m] = 0;
      }
    }
    __syncthreads();
  }

  float acc1[PG_TM][PG_TN], acc2[PG_TM][PG_TN];
#pragma unroll
  for (int i = 0; i < PG_TM; ++i)
#pragma unroll
    for (int j = 0; j < PG_TN; ++j) acc1[i][j] = acc2[i][j] = 0.f;

  const int ak = tid % PG_BK;   // this thread's k within the tile
  const int am = tid / PG_BK;   // first row it loads
  float ra1[A_PER], ra2[A_PER], rb1[B_PER], rb2[B_PER];

  auto load_tiles = [&](int k0) {
    const int kk = k0 + ak;
    if (CONV) {
      int i = 0, j = 0, c = 0;
      int64_t koff = 0;
      const bool kok = kk < K;
      if (kok) {
        int kwC = a.kw * a.C;
        i = kk / kwC;
        int rem = kk - i * kwC;
        j = rem / a.C;
        c = rem - j * a.C;
        koff = ((int64_t)i * a.W + j) * a.C + c;
      }
#pragma unroll
      for (int q = 0; q < A_PER; ++q) {
        int m = am + AROWS * q;
        float v = 0.f;
        int hi = rhi[m] + i, wi = rwi[m] + j;
        if (kok && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W) v = __ldg(a.x + rowoff[m] + koff);
        ra1[q] = v;
        ra2[q] = v * v;
      }
    } else {
#pragma unroll
      for (int q = 0; q < A_PER; ++q) {
        int64_t gm = m0 + am + AROWS * q;
        float v = 0.f, zv = 1.f;
        if (gm < a.M && kk < K) {
          v = __ldg(a.x + gm * K + kk);
          if (a.z) zv = __ldg(a.z + gm * K + kk);
        }
        ra1[q] = v * zv;
        ra2[q] = v * v;
      }
    }
#pragma unroll
    for (int q = 0; q < B_PER; ++q) {
      int e = tid + NT * q;
      int n = e % BN, k = e / BN;
      float w = 0.f, v = 0.f;
      if (k0 + k < K && n0 + n < N) {
        w = __ldg(a.Wm + (int64_t)(k0 + k) * N + n0 + n);
        v = __ldg(a.Vw + (int64_t)(k0 + k) * N + n0 + n);
      }
      rb1[q] = w;
      rb2[q] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int q = 0; q < A_PER; ++q) {
      As1[ak][am + AROWS * q] = ra1[q];
      As2[ak][am + AROWS * q] = ra2[q];
    }
#pragma unroll
    for (int q = 0; q < B_PER; ++q) {
      int e = tid + NT * q;
      Bs1[e / BN][e % BN] = rb1[q];
      Bs2[e / BN][e % BN] = rb2[q];
    }
  };

  load_tiles(0);
  for (int k0 = 0; k0 < K; k0 += PG_BK) {
    store_tiles();
    __syncthreads();
    if (k0 + PG_BK < K) load_tiles(k0 + PG_BK);
#pragma unroll
    for (int k = 0; k < PG_BK; ++k) {
      float4 p0 = *reinterpret_cast<const float4*>(&As1[k][ty * PG_TM]);
      float4 p1 = *reinterpret_cast<const float4*>(&As1[k][ty * PG_TM + 4]);
      float4 q0 = *reinterpret_cast<const float4*>(&As2[k][ty * PG_TM]);
      float4 q1 = *reinterpret_cast<const float4*>(&As2[k][ty * PG_TM + 4]);
      float4 bw = *reinterpret_cast<const float4*>(&Bs1[k][tx * PG_TN]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs2[k][tx * PG_TN]);
      float a1[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      float a2[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      float b1[4] = {bw.x, bw.y, bw.z, bw.w};
      float b2[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < PG_TM; ++i)
#pragma unroll
        for (int j = 0; j < PG_TN; ++j) {
          acc1[i][j] = fmaf(a1[i], b1[j], acc1[i][j]);
          acc2[i][j] = fmaf(a2[i], b2[j], acc2[i][j]);
        }
    }
    __syncthreads();
  }

  // ---- epilogue: bias, (conv) z-scale, sqrt(V) * eps
  const int nb = n0 + tx * PG_TN;
  if (nb >= N) return;
  float bmv[PG_TN], vbv[PG_TN];
#pragma unroll
  for (int j = 0; j < PG_TN; ++j) {
    bmv[j] = nb + j < N ? a.bm[nb + j] : 0.f;
    vbv[j] = nb + j < N ? a.vb[nb + j] : 1.f;
  }
  const int64_t HoWo = CONV ? (int64_t)a.Ho * a.Wo : 1;
#pragma unroll
  for (int i = 0; i < PG_TM; ++i) {
    const int64_t gm = m0 + ty * PG_TM + i;
    if (gm >= a.M) continue;
    const int64_t brow = CONV ? gm / HoWo : gm;
    const uint32_t outer = CONV ? (uint32_t)(gm - brow * HoWo) : 0u;
    float e[4];
    if (a.eps.injected) {
#pragma unroll
      for (int j = 0; j < PG_TN; ++j) e[j] = nb + j < N ? a.eps.injected[gm * N + nb + j] : 0.f;
    } else {
      philox_normal4(a.eps, a.eps.row_offset + (uint64_t)brow, outer, (uint32_t)(nb >> 2), e);
    }
    float out[PG_TN], mo[PG_TN], so[PG_TN];
#pragma unroll
    for (int j = 0; j < PG_TN; ++j) {
      float mean = acc1[i][j] + bmv[j];
      float sd = sqrtf(acc2[i][j] + vbv[j]);
      mo[j] = mean;
      so[j] = sd;
      if (CONV) {
        float zv = (a.z && nb + j < N) ? __ldg(a.z + brow * N + nb + j) : 1.f;
        mean *= zv;
      }
      out[j] = fmaf(sd, e[j], mean);
    }
    float* yp = a.y + gm * N + nb;
    if ((N & 3) == 0) {
      *reinterpret_cast<float4*>(yp) = make_float4(out[0], out[1], out[2], out[3]);
      if (a.sd_out) *reinterpret_cast<float4*>(a.sd_out + gm * N + nb) = make_float4(so[0], so[1], so[2], so[3]);
      if (CONV && a.mean_out)
        *reinterpret_cast<float4*>(a.mean_out + gm * N + nb) = make_float4(mo[0], mo[1], mo[2], mo[3]);
    } else {
#pragma unroll
      for (int j = 0; j < PG_TN; ++j)
        if (nb + j < N) {
          yp[j] = out[j];
          if (a.sd_out) a.sd_out[gm * N + nb + j] = so[j];
          if (CONV && a.mean_out) a.mean_out[gm * N + nb + j] = mo[j];
        }
    }
  }
}

// ----------------------------------------------------------------------- direct small conv
// Few-filter convolutions (LeNet conv1: 5x5x1 -> 20) have a tiny K = kh*kw*C and N = F, so the
// implicit-GEMM tile wastes most of its lanes and the layer is bound by the 46 KB/row output
// write and the eps draws.  Here a CTA stages whole (zero-padded) input images plus both
// filter banks in shared memory; a thread owns TWO horizontally adjacent output pixels and all
// NF filters (2*2*NF accumulators), so every filter word fetched (one broadcast LDS.128 per 4
// filters) feeds 8 FMAs and x is read once per tap.
#define DC_THREADS 256
struct DirectArgs {
  PairedArgs p;
  int Hp, Wp;   // padded input tile
  int pad_b, pad_r;
  int IMG;      // images per CTA
};

template <int NF, bool POOL>
__global__ void __launch_bounds__(DC_THREADS, 2) conv_direct_fwd(DirectArgs d) {
  extern __shared__ __align__(16) float dsm[];
  const PairedArgs& a = d.p;
  const int K = a.K, F = a.N;
  float* wm = dsm;                       // [K][NF]
  float* wv = wm + K * NF;               // [K][NF]
  float* zs = wv + K * NF;               // [IMG][NF]
  int* koff = (int*)(zs + d.IMG * NF);   // [K]
  float* xs = (float*)(koff + ((K + 3) & ~3));  // [IMG][Hp][Wp][C]
  const int tid = threadIdx.x;
  const int img_sz = d.Hp * d.Wp * a.C;
  const int64_t b0 = (int64_t)blockIdx.x * d.IMG;
  const int64_t B = a.M / ((int64_t)a.Ho * a.Wo);
  const int nimg = (int)((B - b0 < d.IMG) ? B - b0 : d.IMG);

  for (int e = tid; e < K * NF; e += DC_THREADS) {
    int k = e / NF, f = e - k * NF;
    wm[e] = f < F ? __ldg(a.Wm + (int64_t)k * F + f) : 0.f;
    wv[e] = f < F ? __ldg(a.Vw + (int64_t)k * F + f) : 0.f;
  }
  for (int e = tid; e < d.IMG * NF; e += DC_THREADS) {
    int im = e / NF, f = e - im * NF;
    zs[e] = (a.z && im < nimg && f < F) ? __ldg(a.z + (b0 + im) * F + f) : 1.f;
  }
  for (int k = tid; k < K; k += DC_THREADS) {
    int kwC = a.kw * a.C;
    int i = k / kwC, rem = k - i * kwC, j = rem / a.C, c = rem - j * a.C;
    koff[k] = (i * d.Wp + j) * a.C + c;
  }
  for (int e = tid; e < d.IMG * img_sz; e += DC_THREADS) {
    int im = e / img_sz, r = e - im * img_sz;
    int c = r % a.C, t = r / a.C;
    int wp = t % d.Wp, hp = t / d.Wp;
    int hi = hp - a.pad_t, wi = wp - a.pad_l;
    float v = 0.f;
    if (im < nimg && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W)
      v = __ldg(a.x + (((b0 + im) * a.H + hi) * a.W + wi) * a.C + c);
    xs[e] = v;
  }
  __syncthreads();

  const int Wo2 = (a.Wo + 1) >> 1;
  const int npairs = nimg * a.Ho * Wo2;
  float bmv[NF], vbv[NF];
#pragma unroll
  for (int f = 0; f < NF; ++f) {
    bmv[f] = f < F ? a.bm[f] : 0.f;
    vbv[f] = f < F ? a.vb[f] : 1.f;
  }
  for (int pbase = 0; pbase < npairs; pbase += DC_THREADS) {
    const int pr0 = pbase + tid;
    const bool valid = pr0 < npairs;
    const int pr = valid ? pr0 : 0;
    int im, ho, wo0;
    if (POOL) {  // lanes l and l^1 own the two rows of one pooling window (Ho, Wo even)
      const int hpar = pr & 1;
      int q = pr >> 1;
      const int wp = q % Wo2;
      q /= Wo2;
      const int ho2 = q % (a.Ho >> 1);
      im = q / (a.Ho >> 1);
      ho = 2 * ho2 + hpar;
      wo0 = 2 * wp;
    } else {
      im = pr / (a.Ho * Wo2);
      const int r = pr - im * a.Ho * Wo2;
      ho = r / Wo2;
      wo0 = (r - ho * Wo2) * 2;
    }
    const bool two = wo0 + 1 < a.Wo;
    const float* x0 = xs + im * img_sz + ((ho * a.stride) * d.Wp + wo0 * a.stride) * a.C;
    const float* x1 = x0 + (two ? a.stride * a.C : 0);
    float m0[NF], v0[NF], m1[NF], v1[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) m0[f] = v0[f] = m1[f] = v1[f] = 0.f;
    for (int k = 0; k < K; ++k) {
      const int o = koff[k];
      const float xa = x0[o], xb = x1[o];
      const float xa2 = xa * xa, xb2 = xb * xb;
      const float4* wmk = reinterpret_cast<const float4*>(wm + k * NF);
      const float4* wvk = reinterpret_cast<const float4*>(wv + k * NF);
#pragma unroll
      for (int q = 0; q < NF / 4; ++q) {
        float4 w = wmk[q], u = wvk[q];
        m0[4 * q + 0] = fmaf(xa, w.x, m0[4 * q + 0]); m0[4 * q + 1] = fmaf(xa, w.y, m0[4 * q + 1]);
        m0[4 * q + 2] = fmaf(xa, w.z, m0[4 * q + 2]); m0[4 * q + 3] = fmaf(xa, w.w, m0[4 * q + 3]);
        m1[4 * q + 0] = fmaf(xb, w.x, m1[4 * q + 0]); m1[4 * q + 1] = fmaf(xb, w.y, m1[4 * q + 1]);
        m1[4 * q + 2] = fmaf(xb, w.z, m1[4 * q + 2]); m1[4 * q + 3] = fmaf(xb, w.w, m1[4 * q + 3]);
        v0[4 * q + 0] = fmaf(xa2, u.x, v0[4 * q + 0]); v0[4 * q + 1] = fmaf(xa2, u.y, v0[4 * q + 1]);
        v0[4 * q + 2] = fmaf(xa2, u.z, v0[4 * q + 2]); v0[4 * q + 3] = fmaf(xa2, u.w, v0[4 * q + 3]);
        v1[4 * q + 0] = fmaf(xb2, u.x, v1[4 * q + 0]); v1[4 * q + 1] = fmaf(xb2, u.y, v1[4 * q + 1]);
        v1[4 * q + 2] = fmaf(xb2, u.z, v1[4 * q + 2]); v1[4 * q + 3] = fmaf(xb2, u.w, v1[4 * q + 3]);
      }
    }
    // epilogue for the one or two pixels
    const int64_t brow = b0 + im;
    if (POOL) {
      const uint32_t pix0 = (uint32_t)(ho * a.Wo + wo0);
#pragma unroll
      for (int q = 0; q < NF / 4; ++q) {
        if (4 * q >= F) break;
        float e0[4], e1[4];
        if (a.eps.injected) {
          const int64_t gm = brow * a.Ho * a.Wo + pix0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            e0[j] = 4 * q + j < F ? a.eps.injected[gm * F + 4 * q + j] : 0.f;
            e1[j] = 4 * q + j < F ? a.eps.injected[(gm + 1) * F + 4 * q + j] : 0.f;
          }
        } else {
          philox_normal4(a.eps, a.eps.row_offset + (uint64_t)brow, pix0, (uint32_t)q, e0);
          philox_normal4(a.eps, a.eps.row_offset + (uint64_t)brow, pix0 + 1, (uint32_t)q, e1);
        }
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int f = 4 * q + j;
          const float zf = zs[im * NF + f];
          const float y0 = fmaf(sqrtf(v0[f] + vbv[f]), e0[j], (m0[f] + bmv[f]) * zf);
          const float y1 = fmaf(sqrtf(v1[f] + vbv[f]), e1[j], (m1[f] + bmv[f]) * zf);
          float v = fmaxf(y0, y1);
          v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));  // the other row of the window
          out[j] = fmaxf(v, 0.f);                            // ReLU
        }
        if (valid && (ho & 1) == 0) {
          const int64_t gp = ((brow * (a.Ho >> 1) + (ho >> 1)) * (a.Wo >> 1) + (wo0 >> 1)) * F + 4 * q;
          if ((F & 3) == 0) {
            *reinterpret_cast<float4*>(a.y + gp) = make_float4(out[0], out[1], out[2], out[3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (4 * q + j < F) a.y[gp + j] = out[j];
          }
        }
      }
      continue;
    }
    if (!valid) continue;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (px == 1 && !two) break;
      const float* mm = px ? m1 : m0;
      const float* vv = px ? v1 : v0;
      const uint32_t pix = (uint32_t)(ho * a.Wo + wo0 + px);
      const int64_t gm = brow * a.Ho * a.Wo + pix;
#pragma unroll
      for (int q = 0; q < NF / 4; ++q) {
        if (4 * q >= F) break;
        float e[4];
        if (a.eps.injected) {
#pragma unroll
          for (int j = 0; j < 4; ++j) e[j] = 4 * q + j < F ? a.eps.injected[gm * F + 4 * q + j] : 0.f;
        } else {
          philox_normal4(a.eps, a.eps.row_offset + (uint64_t)brow, pix, (uint32_t)q, e);
        }
        float out[4], mo[4], so[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int f = 4 * q + j;
          float mean = mm[f] + bmv[f];
          float sd = sqrtf(vv[f] + vbv[f]);
          mo[j] = mean;
          so[j] = sd;
          out[j] = fmaf(sd, e[j], mean * zs[im * NF + f]);
        }
        if ((F & 3) == 0) {
          *reinterpret_cast<float4*>(a.y + gm * F + 4 * q) = make_float4(out[0], out[1], out[2], out[3]);
          if (a.sd_out) *reinterpret_cast<float4*>(a.sd_out + gm * F + 4 * q) = make_float4(so[0], so[1], so[2], so[3]);
          if (a.mean_out) *reinterpret_cast<float4*>(a.mean_out + gm * F + 4 * q) = make_float4(mo[0], mo[1], mo[2], mo[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (4 * q + j < F) {
              a.y[gm * F + 4 * q + j] = out[j];
              if (a.sd_out) a.sd_out[gm * F + 4 * q + j] = so[j];
              if (a.mean_out) a.mean_out[gm * F + 4 * q + j] = mo[j];
            }
        }
      }
    }
  }
}

// returns MNF_OK and sets *used when the direct kernel applies
static int launch_direct(mnf_ctx* ctx, const PairedArgs& a, int pad_b, int pad_r, bool* used) {
  constexpr int NF = 20;
  *used = false;
  if (a.N > NF || a.K > 256) return MNF_OK;
  DirectArgs d;
  d.p = a;
  d.Hp = a.H + a.pad_t + pad_b;
  d.Wp = a.W + a.pad_l + pad_r;
  d.pad_b = pad_b; d.pad_r = pad_r;
  const size_t img_bytes = sizeof(float) * (size_t)d.Hp * d.Wp * a.C;
  const size_t fixed = sizeof(float) * (2 * (size_t)a.K * NF + 8 * NF) + sizeof(int) * ((a.K + 3) & ~3);
  if (fixed + img_bytes > 44 * 1024) return MNF_OK;
  int img = (int)((44 * 1024 - fixed) / img_bytes);
  if (img > 8) img = 8;
  const int64_t B = a.M / ((int64_t)a.Ho * a.Wo);
  d.IMG = img;
  const size_t smem = sizeof(float) * (2 * (size_t)a.K * NF + (size_t)img * NF) + sizeof(int) * ((a.K + 3) & ~3) + img * img_bytes;
  const bool probed = mnf_probe_begin(ctx, MNF_PROBE_CONV_GEMM);
  mnf_note_path("conv_direct");
  if (a.pool)
    conv_direct_fwd<NF, true><<<(unsigned)cdiv(B, img), DC_THREADS, smem, ctx->stream>>>(d);
  else
    conv_direct_fwd<NF, false><<<(unsigned)cdiv(B, img), DC_THREADS, smem, ctx->stream>>>(d);
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  *used = true;
  return MNF_OK;
}

template <bool CONV>
static int launch_paired(mnf_ctx* ctx, const PairedArgs& a) {
  dim3 grid((unsigned)cdiv(a.M, PG_BM), 1);
  const bool probed = mnf_probe_begin(ctx, CONV ? MNF_PROBE_CONV_GEMM : MNF_PROBE_DENSE_GEMM);
  mnf_note_path("paired_gemm");
  if (a.N > 32) {
    grid.y = (unsigned)cdiv(a.N, 64);
    paired_gemm_fwd<CONV, 64><<<grid, (PG_BM / PG_TM) * (64 / PG_TN), 0, ctx->stream>>>(a);
  } else if (a.N > 16) {
    paired_gemm_fwd<CONV, 32><<<grid, (PG_BM / PG_TM) * (32 / PG_TN), 0, ctx->stream>>>(a);
  } else {
    paired_gemm_fwd<CONV, 16><<<grid, (PG_BM / PG_TM) * (16 / PG_TN), 0, ctx->stream>>>(a);
  }
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

// ------------------------------------------------------------------ dense layers with few outputs
// MNFDense.call (mnf_dense.py:159-170) for N <= 16 (the classifier head: 500 -> 10).  A 128-row
// tensor-core tile would waste 5/6 of its columns; here a warp owns a row: lanes split K (float4
// along k, weights transposed in shared memory so that the reads are conflict-free), 2N partial
// sums per lane, butterfly reduction, lane n finishes output n: bias, Philox eps, optional ReLU or
// Softmax (warp-wide, same formula as softmax_kernel) -- one launch instead of GEMM + activation.
struct SmallDenseArgs {
  int64_t B;
  int K, N, act;  // act: 0 none, 1 ReLU, 2 Softmax
  const float *x, *z, *Wm, *Ls, *bm, *lvb;
  float max_std;
  float *y, *sd_out;
  NoiseDev eps;
};
constexpr int SD_NMAX = 16;

// NT: N rounded up to a multiple of 4 (compile-time loop bounds; the padding weights are zero)
template <int NT>
__global__ void __launch_bounds__(256) dense_small_fwd(SmallDenseArgs a) {
  extern __shared__ __align__(16) float sd_smem[];
  const int K4 = a.K >> 2, Kp = a.K + 4;        // transposed rows padded by one float4
  float* Wt = sd_smem;                           // [N][Kp]  mean weights
  float* Vt = Wt + NT * Kp;                      // [NT][Kp] clipped variances
  for (int i = threadIdx.x; i < (NT - a.N) * Kp; i += blockDim.x) {  // zero rows of the padding outputs
    Wt[a.N * Kp + i] = 0.f;
    Vt[a.N * Kp + i] = 0.f;
  }
  {
    const int tot = a.K * a.N;
    for (int i0 = threadIdx.x; i0 < tot; i0 += 4 * blockDim.x) {  // four independent loads in flight
      float w[4], l[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        w[u] = i < tot ? __ldg(a.Wm + i) : 0.f;
        l[u] = i < tot ? __ldg(a.Ls + i) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < tot) {
          const int k = i / a.N, n = i - k * a.N;
          Wt[n * Kp + k] = w[u];
          const float sd = fminf(expf(l[u]), a.max_std);
          Vt[n * Kp + k] = sd * sd;
        }
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float bmn = lane < a.N ? a.bm[lane] : 0.f;
  const float vbn = lane < a.N ? fminf(expf(a.lvb[lane]), a.max_std * a.max_std) : 1.f;
  constexpr int UI = 4;  // k4 iterations whose loads are issued together (K <= 512 in one go)
  const int64_t rstride = (int64_t)gridDim.x * wpb;
  for (int64_t row = blockIdx.x * (int64_t)wpb + warp; row < a.B; row += rstride) {
    float am[NT], av[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) am[n] = av[n] = 0.f;
    const float4* xr = reinterpret_cast<const float4*>(a.x + row * a.K);
    const float4* zr = a.z ? reinterpret_cast<const float4*>(a.z + row * a.K) : nullptr;
    for (int kb = 0; kb < K4; kb += 32 * UI) {
      float4 xv[UI], zv[UI];
#pragma unroll
      for (int u = 0; u < UI; ++u) {
        const int k4 = kb + lane + 32 * u;
        xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        zv[u] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (k4 < K4) {
          xv[u] = __ldg(xr + k4);
          if (zr) zv[u] = __ldg(zr + k4);
        }
      }
#pragma unroll
      for (int u = 0; u < UI; ++u) {
        const int k4 = kb + lane + 32 * u;
        if (k4 < K4) {
          const float xz[4] = {xv[u].x * zv[u].x, xv[u].y * zv[u].y, xv[u].z * zv[u].z, xv[u].w * zv[u].w};
          const float x2[4] = {xv[u].x * xv[u].x, xv[u].y * xv[u].y, xv[u].z * xv[u].z, xv[u].w * xv[u].w};
          const float* wp = Wt + 4 * k4;
          const float* vp = Vt + 4 * k4;
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const float4 w = *reinterpret_cast<const float4*>(wp + n * Kp);
            const float4 v = *reinterpret_cast<const float4*>(vp + n * Kp);
            am[n] = fmaf(xz[0], w.x, fmaf(xz[1], w.y, fmaf(xz[2], w.z, fmaf(xz[3], w.w, am[n]))));
            av[n] = fmaf(x2[0], v.x, fmaf(x2[1], v.y, fmaf(x2[2], v.z, fmaf(x2[3], v.w, av[n]))));
          }
        }
      }
    }
    float mean = 0.f, var = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const float m = warp_sum(am[n]), v = warp_sum(av[n]);
      if (n == lane) { mean = m; var = v; }
    }
    float out = 0.f;
    if (lane < a.N) {
      const float sd = sqrtf(var + vbn);
      if (a.sd_out) a.sd_out[row * a.N + lane] = sd;
      out = fmaf(sd, noise_normal1(a.eps, row, 0u, (uint32_t)lane, 1, a.N), mean + bmn);
      if (a.act == 1) out = fmaxf(out, 0.f);
    }
    if (a.act == 2) {  // warp-wide softmax over the N outputs (stock.cu:softmax_kernel arithmetic)
      float mx = lane < a.N ? out : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float e = lane < a.N ? expf(out - mx) : 0.f;
      const float ssum = warp_sum(e);
      out = e / ssum;
    }
    if (lane < a.N) a.y[row * a.N + lane] = out;
  }
}

// The same layer for MANY rows (MNF-LeNet's 500 -> 10 head at 10240 rows).  The warp-per-row kernel above is bound
// by its shared-memory WEIGHT reads: 2 NT float4 loads per k-quad and row, and a 128-bit warp-wide shared load
// costs its four quarter-warp phases whether or not the lanes' addresses coincide (measured: sharing a load
// between the lanes of different rows changed nothing, 41 us either way).  What helps is re-using a loaded weight
// quad from REGISTERS: here a lane owns RPL rows (eight lanes split K per row group, four row groups per warp), so
// every weight / variance float4 feeds RPL rows' FMAs; three shuffle steps finish the contraction and lane j < 4
// of a group completes outputs 4j .. 4j+3 of its rows with one Philox call per row.
// [2][NT][K + 4]: transposed mean weights, then clipped variances (zero rows for the padding outputs); packed once
// per call -- or once per weight set when frozen -- so that a CTA's prologue is ONE bulk copy instead of ten
// dependent rounds of scattered loads (which cost more than the contraction itself)
__global__ void small_pack_weights(int K, int N, int NT, const float* __restrict__ Wm, const float* __restrict__ Ls, float max_std,
                                   float* __restrict__ WV) {
  const int Kp = K + 4, tot = NT * Kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += gridDim.x * blockDim.x) {
    const int n = i / Kp, k = i - n * Kp;
    float w = 0.f, v = 0.f;
    if (n < N && k < K) {
      w = Wm[(int64_t)k * N + n];
      const float sd = fminf(expf(Ls[(int64_t)k * N + n]), max_std);
      v = sd * sd;
    }
    WV[i] = w;
    WV[tot + i] = v;
  }
}
template <int NT, int RPL>
__global__ void __launch_bounds__(128) dense_small_rows_fwd(SmallDenseArgs a, const float* __restrict__ WV) {
  extern __shared__ __align__(128) float sd_smem[];
  const int K4 = a.K >> 2, Kp = a.K + 4;
  float* Wt = sd_smem;
  float* Vt = Wt + NT * Kp;
  uint64_t* wbar = reinterpret_cast<uint64_t*>(Vt + NT * Kp);
  const uint32_t wbar_s = (uint32_t)__cvta_generic_to_shared(wbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wbar_s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = 2u * NT * Kp * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wbar_s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(Wt)),
                 "l"(WV), "r"(bytes), "r"(wbar_s)
                 : "memory");
  }
  __syncthreads();
  {
    uint32_t done = 0;
    for (uint32_t polls = 0; !done; ++polls) {  // bounded: a copy that never lands ends as a launch failure
      asm volatile(
          "{\n\t.reg .pred P1;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, P1;\n\t}"
          : "=r"(done)
          : "r"(wbar_s)
          : "memory");
      if (!done && polls > (1u << 24)) __trap();
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int rg = lane >> 3, j8 = lane & 7, j = j8 & 3;  // (lanes 4-7 of a group end up with copies and do not store)
  float bm4[4], vb4[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int n = 4 * j + e;
    bm4[e] = n < a.N ? a.bm[n] : 0.f;
    vb4[e] = n < a.N ? fminf(expf(a.lvb[n]), a.max_std * a.max_std) : 1.f;
  }
  constexpr int RW = 4 * RPL;  // rows per warp
  const int64_t rstride = (int64_t)gridDim.x * wpb * RW;
  for (int64_t rw0 = (blockIdx.x * (int64_t)wpb + warp) * RW; rw0 < a.B; rw0 += rstride) {
    const int64_t row0 = rw0 + rg * RPL;  // this lane group's first row
    float am[RPL][NT], av[RPL][NT];
    const float4* xr[RPL];
    const float4* zr[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
#pragma unroll
      for (int n = 0; n < NT; ++n) am[r][n] = av[r][n] = 0.f;
      const int64_t rowc = row0 + r < a.B ? row0 + r : a.B - 1;
      xr[r] = reinterpret_cast<const float4*>(a.x + rowc * a.K);
      zr[r] = a.z ? reinterpret_cast<const float4*>(a.z + rowc * a.K) : nullptr;
    }
    constexpr int UI = 4;  // k-quads per lane whose loads are issued together
    for (int kb = 0; kb < K4; kb += 8 * UI) {
      float4 xv[UI][RPL], zv[UI][RPL];
#pragma unroll
      for (int u = 0; u < UI; ++u) {
        const int k4 = kb + j8 + 8 * u;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
          xv[u][r] = make_float4(0.f, 0.f, 0.f, 0.f);
          zv[u][r] = make_float4(1.f, 1.f, 1.f, 1.f);
          if (k4 < K4) {
            xv[u][r] = __ldg(xr[r] + k4);
            if (zr[r]) zv[u][r] = __ldg(zr[r] + k4);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UI; ++u) {
        const int k4 = kb + j8 + 8 * u;
        if (k4 < K4) {
          float xz[RPL][4], x2[RPL][4];
#pragma unroll
          for (int r = 0; r < RPL; ++r) {
            xz[r][0] = xv[u][r].x * zv[u][r].x; xz[r][1] = xv[u][r].y * zv[u][r].y;
            xz[r][2] = xv[u][r].z * zv[u][r].z; xz[r][3] = xv[u][r].w * zv[u][r].w;
            x2[r][0] = xv[u][r].x * xv[u][r].x; x2[r][1] = xv[u][r].y * xv[u][r].y;
            x2[r][2] = xv[u][r].z * xv[u][r].z; x2[r][3] = xv[u][r].w * xv[u][r].w;
          }
          const float* wp = Wt + 4 * k4;
          const float* vp = Vt + 4 * k4;
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            const float4 w = *reinterpret_cast<const float4*>(wp + n * Kp);
            const float4 v = *reinterpret_cast<const float4*>(vp + n * Kp);
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
              am[r][n] = fmaf(xz[r][0], w.x, fmaf(xz[r][1], w.y, fmaf(xz[r][2], w.z, fmaf(xz[r][3], w.w, am[r][n]))));
              av[r][n] = fmaf(x2[r][0], v.x, fmaf(x2[r][1], v.y, fmaf(x2[r][2], v.z, fmaf(x2[r][3], v.w, av[r][n]))));
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
#pragma unroll
      for (int n = 0; n < NT; ++n) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          am[r][n] += __shfl_xor_sync(0xffffffffu, am[r][n], o);
          av[r][n] += __shfl_xor_sync(0xffffffffu, av[r][n], o);
        }
      }
      const int64_t row = row0 + r;
      const bool row_ok = row < a.B;
      const int64_t rowc = row_ok ? row : a.B - 1;
      // lane j of the group: outputs 4j .. 4j+3 of this row
      float m4[4], v4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        m4[e] = am[r][e]; v4[e] = av[r][e];
        if (NT > 4 && j == 1) { m4[e] = am[r][(4 + e) % NT]; v4[e] = av[r][(4 + e) % NT]; }
        if (NT > 8 && j == 2) { m4[e] = am[r][(8 + e) % NT]; v4[e] = av[r][(8 + e) % NT]; }
        if (NT > 12 && j == 3) { m4[e] = am[r][(12 + e) % NT]; v4[e] = av[r][(12 + e) % NT]; }
      }
      float e4[4] = {0.f, 0.f, 0.f, 0.f};
      if (4 * j < a.N) {
        if (a.eps.injected) {
#pragma unroll
          for (int e = 0; e < 4; ++e) e4[e] = 4 * j + e < a.N ? a.eps.injected[rowc * a.N + 4 * j + e] : 0.f;
        } else {
          philox_normal4(a.eps, a.eps.row_offset + (uint64_t)rowc, 0u, (uint32_t)j, e4);
        }
      }
      float out[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = 4 * j + e;
        const float sd = sqrtf(v4[e] + vb4[e]);
        if (a.sd_out && row_ok && j8 < 4 && n < a.N) a.sd_out[row * a.N + n] = sd;
        out[e] = fmaf(sd, e4[e], m4[e] + bm4[e]);
        if (a.act == 1) out[e] = fmaxf(out[e], 0.f);
      }
      if (a.act == 2) {  // softmax over the row's N outputs (lanes 0-3 of the group; stock.cu:softmax_kernel arithmetic)
        float mx = -INFINITY;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * j + e < a.N) mx = fmaxf(mx, out[e]);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float ssum = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          out[e] = 4 * j + e < a.N ? expf(out[e] - mx) : 0.f;
          ssum += out[e];
        }
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
        ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
#pragma unroll
        for (int e = 0; e < 4; ++e) out[e] = out[e] / ssum;
      }
      if (row_ok && j8 < 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * j + e < a.N) a.y[row * a.N + 4 * j + e] = out[e];
      }
    }
  }
}

static bool small_dense_ok(int64_t K, int64_t N, const float* x, const float* z) {
  return N <= SD_NMAX && (K & 3) == 0 && K * ((N + 3) & ~3) <= 12288 && (((uintptr_t)x | (uintptr_t)z) & 15) == 0;
}

// workspace layout for one layer call: Vw [K*N] then vb [N]
static int prep_var(mnf_ctx* ctx, int64_t KN, int N, const float* Ls, const float* lvb, float max_std, float** Vw,
                    float** vb) {
  size_t bytes = sizeof(float) * (size_t)(KN + N + 8);
  float* ws = (float*)mnf_workspace(ctx, bytes);
  if (!ws) return MNF_ERR_CUDA;
  *Vw = ws;
  *vb = ws + ((KN + 3) / 4) * 4;
  int64_t nthr = KN > N ? KN : N;
  int grid = (int)(cdiv(nthr, 256) < ctx->sm_count * 8 ? cdiv(nthr, 256) : ctx->sm_count * 8);
  prep_var_kernel<<<grid, 256, 0, ctx->stream>>>(nthr, Ls, max_std, *Vw, N, lvb, *vb);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

// ---- long contractions in the fp32-parity tensor-core mode (K > MNF_TC_PARITY_KMAX): the fused
// tcgen05 kernel would sum K terms in one TMEM accumulator (drift beyond the 1e-5 bar) and the FFMA
// kernel is ~8x slower, so the two products run as split-K GEMMs (gemm_umma.cu: chains of <= 1024
// terms combined by fp32 atomics) and a small kernel applies the local-reparameterisation epilogue.
__global__ void dense_xz_kernel(int64_t n, const float* __restrict__ x, const float* __restrict__ z, float* __restrict__ A) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) A[i] = x[i] * z[i];
}
__global__ void dense_long_epilogue(int64_t B, int N, const float* __restrict__ mean, const float* __restrict__ var,
                                    const float* __restrict__ bm, const float* __restrict__ vb, NoiseDev eps, int act,
                                    float* __restrict__ y, float* __restrict__ sd_out) {
  const int N4 = (N + 3) >> 2;
  const int64_t total = B * N4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = q / N4;
    const int nb = (int)(q - m * N4) * 4;
    float e[4];
    if (eps.injected) {
#pragma unroll
      for (int j = 0; j < 4; ++j) e[j] = nb + j < N ? eps.injected[m * N + nb + j] : 0.f;
    } else {
      philox_normal4(eps, eps.row_offset + (uint64_t)m, 0u, (uint32_t)(nb >> 2), e);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = nb + j;
      if (n < N) {
        const float sd = sqrtf(var[m * N + n] + vb[n]);
        const float out = fmaf(sd, e[j], mean[m * N + n] + bm[n]);
        y[m * N + n] = act == MNF_ACT_RELU ? fmaxf(out, 0.f) : out;
        if (sd_out) sd_out[m * N + n] = sd;
      }
    }
  }
}
static int dense_long_forward(mnf_ctx* ctx, int64_t B, int64_t K, int64_t N, const float* x, const float* z,
                              const float* mean_W, const float* log_std_W, const float* mean_b, const float* log_var_b,
                              float max_std, const NoiseDev& eps, int act, float* y, float* sd_out) {
  const size_t pv_floats = (((size_t)(K * N + N + 8)) + 63) & ~(size_t)63;  // prep_var's block comes first
  const size_t BK = (size_t)B * K, BN = (size_t)B * N;
  float* ws = (float*)mnf_workspace(ctx, sizeof(float) * (pv_floats + (z ? BK : 0) + 2 * BN + 64));
  if (!ws) return MNF_ERR_CUDA;
  float *Vw, *vb;
  int rc = prep_var(ctx, K * N, (int)N, log_std_W, log_var_b, max_std, &Vw, &vb);  // same base pointer: no growth
  if (rc) return rc;
  float* A = ws + pv_floats;
  float* mean = A + (z ? BK : 0);
  float* var = mean + BN;
  const float* Aop = x;
  const int cap = ctx->sm_count * 8;
  if (z) {
    const int64_t g = cdiv((int64_t)BK, 256);
    dense_xz_kernel<<<(int)(g < cap ? g : cap), 256, 0, ctx->stream>>>((int64_t)BK, x, z, A);
    MNF_LAUNCHED(ctx);
    Aop = A;
  }
  const bool probed = mnf_probe_begin(ctx, MNF_PROBE_DENSE_GEMM);  // brackets both GEMMs and the epilogue
  if ((rc = mnf_gemm_any(ctx, (int)B, (int)N, (int)K, Aop, K, 1, mean_W, N, 1, mean, N, false, false))) return rc;
  if ((rc = mnf_gemm_any(ctx, (int)B, (int)N, (int)K, x, K, 1, Vw, N, 1, var, N, false, true))) return rc;
  const int64_t g = cdiv(B * ((N + 3) / 4), 256);
  mnf_note_path("dense_long_gemm");
  dense_long_epilogue<<<(int)(g < cap ? g : cap), 256, 0, ctx->stream>>>(B, (int)N, mean, var, mean_b, vb, eps, act, y, sd_out);
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

extern "C" {

int mnf_conv_out_hw(const mnf_conv_shape* s, int64_t* Ho, int64_t* Wo) {
  MNF_CHECK_ARG(s && Ho && Wo, "mnf_conv_out_hw: NULL argument");
  MNF_CHECK_ARG(s->stride >= 1 && s->kh >= 1 && s->kw >= 1, "mnf_conv_out_hw: bad kernel/stride");
  if (s->same) {
    *Ho = cdiv(s->H, s->stride);
    *Wo = cdiv(s->W, s->stride);
  } else {
    MNF_CHECK_ARG(s->H >= s->kh && s->W >= s->kw, "conv VALID: input %lldx%lld smaller than kernel %lldx%lld",
                  (long long)s->H, (long long)s->W, (long long)s->kh, (long long)s->kw);
    *Ho = (s->H - s->kh) / s->stride + 1;
    *Wo = (s->W - s->kw) / s->stride + 1;
  }
  return MNF_OK;
}

int mnf_conv2d_can_fuse_pool(mnf_ctx* ctx, const mnf_conv_shape* s, int32_t* ok) {
  MNF_CHECK_ARG(ctx && s && ok, "mnf_conv2d_can_fuse_pool: NULL argument");
  *ok = 0;
  int64_t Ho, Wo;
  int rc = mnf_conv_out_hw(s, &Ho, &Wo);
  if (rc) return rc;
  if ((Ho & 1) || (Wo & 1)) return MNF_OK;
  const int64_t Kc = s->kh * s->kw * s->C;
  if (s->F <= 20 && Kc <= 256) {  // direct small-filter kernel (same smem test as launch_direct)
    int64_t ph = 0, pw = 0;
    if (s->same) {
      ph = (Ho - 1) * s->stride + s->kh - s->H;
      pw = (Wo - 1) * s->stride + s->kw - s->W;
      ph = ph > 0 ? ph : 0;
      pw = pw > 0 ? pw : 0;
    }
    const size_t img_bytes = sizeof(float) * (size_t)((s->H + ph) * (s->W + pw) * s->C);
    const size_t fixed = sizeof(float) * (2 * (size_t)Kc * 20 + 8 * 20) + sizeof(int) * ((Kc + 3) & ~3);
    if (fixed + img_bytes <= 44 * 1024) {
      *ok = 1;
      return MNF_OK;
    }
  }
  if (mnf_tc_k_ok(ctx, (int64_t)s->kh * s->kw * s->C) && s->F <= 64 && s->F > 20 && 128 % (2 * Wo) == 0) *ok = 1;
  return MNF_OK;
}

// which activations the forward kernel of this shape / math mode can apply in its epilogue
static bool dense_act_ok(const mnf_ctx* ctx, int64_t K, int64_t N, const float* x, const float* z, int act) {
  if (act == MNF_ACT_NONE) return true;
  if (small_dense_ok(K, N, x, z)) return act == MNF_ACT_RELU || act == MNF_ACT_SOFTMAX;
  return act == MNF_ACT_RELU && ctx->math_mode != MNF_MATH_FP32;
}

int mnf_dense_can_fuse_act(mnf_ctx* ctx, int64_t K, int64_t N, int32_t act, int32_t* ok) {
  MNF_CHECK_ARG(ctx && ok && K >= 1 && N >= 1, "mnf_dense_can_fuse_act: bad argument");
  MNF_CHECK_ARG(act >= MNF_ACT_NONE && act <= MNF_ACT_SOFTMAX, "mnf_dense_can_fuse_act: unknown activation %d", act);
  *ok = dense_act_ok(ctx, K, N, nullptr, nullptr, act) ? 1 : 0;  // buffers from this library are 256-byte aligned
  return MNF_OK;
}

int mnf_dense_forward_act(mnf_ctx* ctx, int64_t B, int64_t K, int64_t N, const float* x, const float* z,
                          const float* mean_W, const float* log_std_W, const float* mean_b, const float* log_var_b,
                          float max_std, const mnf_noise* eps, int32_t act, float* y, float* sd_out) {
  MNF_CHECK_ARG(ctx && x && mean_W && log_std_W && mean_b && log_var_b && y, "mnf_dense_forward: NULL argument");
  MNF_CHECK_ARG(B >= 0 && K >= 1 && N >= 1 && K < (1 << 30) && N < (1 << 30), "mnf_dense_forward: bad shape");
  MNF_CHECK_ARG(eps, "mnf_dense_forward: eps noise spec is NULL");
  MNF_CHECK_ARG(act >= MNF_ACT_NONE && act <= MNF_ACT_SOFTMAX, "mnf_dense_forward: unknown activation %d", act);
  if (!dense_act_ok(ctx, K, N, x, z, act)) {
    mnf_set_error("mnf_dense_forward: activation %d cannot be fused for K=%lld N=%lld in this math mode", act,
                  (long long)K, (long long)N);
    return MNF_ERR_UNSUPPORTED;
  }
  if (B == 0) return MNF_OK;
  if (small_dense_ok(K, N, x, z)) {
    SmallDenseArgs a;
    a.B = B; a.K = (int)K; a.N = (int)N; a.act = act;
    a.x = x; a.z = z; a.Wm = mean_W; a.Ls = log_std_W; a.bm = mean_b; a.lvb = log_var_b;
    a.max_std = max_std; a.y = y; a.sd_out = sd_out; a.eps = make_noise(eps);
    const int NT = ((int)N + 3) & ~3;
    const size_t smem = sizeof(float) * 2 * (size_t)NT * (K + 4);
    auto kern = NT == 4 ? dense_small_fwd<4> : NT == 8 ? dense_small_fwd<8> : NT == 12 ? dense_small_fwd<12> : dense_small_fwd<16>;
    static size_t attr_smem[4] = {0, 0, 0, 0};
    if (smem > attr_smem[NT / 4 - 1]) {
      MNF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem[NT / 4 - 1] = smem;
    }
    if (B >= 1024) {  // many rows: every weight quad read from shared memory feeds RPL rows
      constexpr int RPL = 2;
      auto k8 = NT == 4 ? dense_small_rows_fwd<4, RPL> : NT == 8 ? dense_small_rows_fwd<8, RPL> : NT == 12 ? dense_small_rows_fwd<12, RPL>
                                                                                                  : dense_small_rows_fwd<16, RPL>;
      static size_t attr8[4] = {0, 0, 0, 0};
      if (smem > attr8[NT / 4 - 1]) {
        MNF_CUDA(cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 16));
        attr8[NT / 4 - 1] = smem;
      }
      int per_sm8 = 1;
      MNF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm8, k8, 128, smem + 16));
      per_sm8 = per_sm8 < 1 ? 1 : per_sm8;
      const int64_t blocks8 = cdiv(B, 16 * RPL), cap8 = (int64_t)ctx->sm_count * per_sm8;
      bool fresh = true;
      float* WV = (float*)mnf_packed(ctx, mnf_mix(0x5DA11ull, ((uint64_t)K << 32) | (uint32_t)N), (uint64_t)(uintptr_t)mean_W,
                                     (uint64_t)(uintptr_t)log_std_W, (uint64_t)__float_as_uint_host(max_std),
                                     sizeof(float) * 2 * (size_t)NT * (K + 4), &fresh);
      if (!WV) return MNF_ERR_CUDA;
      if (fresh) {
        small_pack_weights<<<(unsigned)cdiv((int64_t)NT * (K + 4), 256), 256, 0, ctx->stream>>>((int)K, (int)N, NT, mean_W, log_std_W, max_std, WV);
        MNF_LAUNCHED(ctx);
      }
      mnf_note_path("dense_small_rows");
      const bool probed8 = mnf_probe_begin(ctx, MNF_PROBE_DENSE_GEMM);
      k8<<<(int)(blocks8 < cap8 ? blocks8 : cap8), 128, smem + 16, ctx->stream>>>(a, WV);
      mnf_probe_end(ctx, probed8);
      MNF_LAUNCHED(ctx);
      return MNF_OK;
    }
    // latency-bound per warp (load -> FMAs -> shuffles -> Philox): fill the SMs with warps, as many
    // CTAs per SM as the transposed weights in shared memory allow
    const int64_t blocks = cdiv(B, 8);
    int per_sm = 1;
    MNF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    per_sm = per_sm < 1 ? 1 : per_sm;  // one wave of resident CTAs; rows are grid-strided
    const int64_t cap = (int64_t)ctx->sm_count * per_sm;
    const int grid = (int)(blocks < cap ? blocks : cap);
    mnf_note_path("dense_small");
    const bool probed = mnf_probe_begin(ctx, MNF_PROBE_DENSE_GEMM);
    kern<<<grid, 256, smem, ctx->stream>>>(a);
    mnf_probe_end(ctx, probed);
    MNF_LAUNCHED(ctx);
    return MNF_OK;
  }
  if (ctx->math_mode == MNF_MATH_BF16)
    return mnf_dense_bf16_forward(ctx, B, (int)K, (int)N, x, z, mean_W, log_std_W, mean_b, log_var_b, max_std,
                                  make_noise(eps), y, sd_out, act);
  // long contractions of MANY rows in the fp32-parity mode: the packed fp16 GEMM, one launch per 1024-term segment
  if (!mnf_tc_k_ok(ctx, K) && mnf_dense_f16x3_ok(ctx, B, (int)K, (int)N) && !getenv("MNF_NO_F16X3"))
    return mnf_dense_f16x3_forward(ctx, B, (int)K, (int)N, x, z, mean_W, log_std_W, mean_b, log_var_b, max_std,
                                   make_noise(eps), y, sd_out, act);
  if (ctx->math_mode == MNF_MATH_TF32X3 && !mnf_tc_k_ok(ctx, K) && B < (1ll << 31))
    return dense_long_forward(ctx, B, K, N, x, z, mean_W, log_std_W, mean_b, log_var_b, max_std, make_noise(eps), act, y,
                              sd_out);
  // many rows: operands packed once into split fp16 planes, GEMM fed by bulk copies (dense_f16x3.cu);
  // few rows (training minibatches): the one-launch tile kernel whose producer warps split in place
  if (mnf_dense_f16x3_ok(ctx, B, (int)K, (int)N) && !getenv("MNF_NO_F16X3"))
    return mnf_dense_f16x3_forward(ctx, B, (int)K, (int)N, x, z, mean_W, log_std_W, mean_b, log_var_b, max_std,
                                   make_noise(eps), y, sd_out, act);
  if (ctx->math_mode == MNF_MATH_TF32X3 && mnf_tc_k_ok(ctx, K))
    return mnf_tc_paired_forward(ctx, 0, B, (int)K, (int)N, x, z, mean_W, log_std_W, mean_b, log_var_b, max_std,
                                 make_noise(eps), y, nullptr, sd_out, nullptr, act);
  float *Vw, *vb;
  // note: the second rounded-up offset inside prep_var must match (K*N padded to 4)
  int rc = prep_var(ctx, K * N, (int)N, log_std_W, log_var_b, max_std, &Vw, &vb);
  if (rc) return rc;
  PairedArgs a = {};
  a.M = B; a.K = (int)K; a.N = (int)N;
  a.x = x; a.z = z; a.Wm = mean_W; a.Vw = Vw; a.bm = mean_b; a.vb = vb;
  a.y = y; a.mean_out = nullptr; a.sd_out = sd_out; a.eps = make_noise(eps);
  return launch_paired<false>(ctx, a);
}

int mnf_dense_forward(mnf_ctx* ctx, int64_t B, int64_t K, int64_t N, const float* x, const float* z,
                      const float* mean_W, const float* log_std_W, const float* mean_b, const float* log_var_b,
                      float max_std, const mnf_noise* eps, float* y, float* sd_out) {
  return mnf_dense_forward_act(ctx, B, K, N, x, z, mean_W, log_std_W, mean_b, log_var_b, max_std, eps, MNF_ACT_NONE, y,
                               sd_out);
}

int mnf_conv2d_forward(mnf_ctx* ctx, const mnf_conv_shape* s, const float* x, const float* z, const float* mean_W,
                       const float* log_std_W, const float* mean_b, const float* log_var_b, float max_std,
                       const mnf_noise* eps, float* y, float* mean_out, float* sd_out) {
  MNF_CHECK_ARG(ctx && s && x && mean_W && log_std_W && mean_b && log_var_b && y, "mnf_conv2d_forward: NULL argument");
  MNF_CHECK_ARG(eps, "mnf_conv2d_forward: eps noise spec is NULL");
  MNF_CHECK_ARG(s->B >= 0 && s->H >= 1 && s->W >= 1 && s->C >= 1 && s->F >= 1 && s->H < 16384 && s->W < 16384,
                "mnf_conv2d_forward: bad input shape");
  int64_t Ho, Wo;
  int rc = mnf_conv_out_hw(s, &Ho, &Wo);
  if (rc) return rc;
  if (s->B == 0) return MNF_OK;
  const int64_t Kc = s->kh * s->kw * s->C;
  MNF_CHECK_ARG(Kc < (1 << 30) && s->F < (1 << 30), "mnf_conv2d_forward: filter too large");
  PairedArgs a = {};
  a.M = s->B * Ho * Wo; a.K = (int)Kc; a.N = (int)s->F;
  a.x = x; a.z = z; a.Wm = mean_W; a.Vw = nullptr; a.bm = mean_b; a.vb = nullptr;
  auto need_var = [&]() -> int {  // only the FFMA kernels read the clipped variances from global memory
    if (a.Vw) return MNF_OK;
    float *Vw, *vb;
    const int r = prep_var(ctx, Kc * s->F, (int)s->F, log_std_W, log_var_b, max_std, &Vw, &vb);
    a.Vw = Vw; a.vb = vb;
    return r;
  };
  a.y = y; a.mean_out = mean_out; a.sd_out = sd_out; a.eps = make_noise(eps);
  a.H = (int)s->H; a.W = (int)s->W; a.C = (int)s->C; a.kh = (int)s->kh; a.kw = (int)s->kw;
  a.Ho = (int)Ho; a.Wo = (int)Wo; a.stride = (int)s->stride;
  a.pad_t = a.pad_l = 0;
  int pad_b = 0, pad_r = 0;
  if (s->same) {  // TF SAME: total pad = max((out-1)*stride + k - in, 0), smaller half first
    int64_t ph = (Ho - 1) * s->stride + s->kh - s->H, pw = (Wo - 1) * s->stride + s->kw - s->W;
    ph = ph > 0 ? ph : 0;
    pw = pw > 0 ? pw : 0;
    a.pad_t = (int)(ph / 2);
    a.pad_l = (int)(pw / 2);
    pad_b = (int)(ph - ph / 2);
    pad_r = (int)(pw - pw / 2);
  }
  a.pool = s->fuse_relu_pool ? 1 : 0;
  if (a.pool) {
    int32_t ok = 0;
    rc = mnf_conv2d_can_fuse_pool(ctx, s, &ok);
    if (rc) return rc;
    if (!ok || mean_out || sd_out) {
      mnf_set_error("mnf_conv2d_forward: fuse_relu_pool is not available for this shape / math mode / training call");
      return MNF_ERR_UNSUPPORTED;
    }
  }
  bool used = false;
  const int geo[11] = {a.H, a.W, a.C, a.kh, a.kw, a.Ho, a.Wo, a.stride, a.pad_t, a.pad_l, a.pool};
  if (ctx->math_mode != MNF_MATH_FP32 && !pad_b && !pad_r && !mean_out && !sd_out) {
    // single-channel input (MNF-LeNet conv1): sliding-window planes on tcgen05
    rc = mnf_conv_win_umma_forward(ctx, s->B, geo, a.N, x, z, mean_W, log_std_W, mean_b, log_var_b, max_std, a.eps, y, &used);
    if (rc || used) return rc;
  }
  const bool tc_ok = mnf_tc_k_ok(ctx, a.K);  // long contractions stay on FFMA in the fp32-parity mode
  // shapes of the direct FFMA kernel; a training call on the tensor-core modes (mean / sd saved for the
  // backward pass) takes the implicit GEMM instead: the direct kernel is tuned for inference row counts
  if (!tc_ok || (a.N <= 20 && a.K <= 256 && !(tc_ok && (mean_out || sd_out)))) {
    rc = need_var();
    if (rc) return rc;
    rc = launch_direct(ctx, a, pad_b, pad_r, &used);
    if (rc || used) return rc;
  }
  if (tc_ok) {
    if (!pad_b && !pad_r) {  // raw-tile kernel: im2col expressed by UMMA descriptors (8-wide output rows)
      if (!getenv("MNF_NO_F16")) {  // scaled fp16 hi/lo operands: half the weight stream and MMAs
        rc = mnf_conv_raw_f16_forward(ctx, s->B, geo, a.N, x, z, mean_W, log_std_W, mean_b, log_var_b, max_std, a.eps, y,
                                      mean_out, sd_out, &used);
        if (rc || used) return rc;
      }
      rc = mnf_conv_raw_umma_forward(ctx, s->B, geo, a.N, x, z, mean_W, log_std_W, mean_b, log_var_b, max_std, a.eps, y,
                                     mean_out, sd_out, &used);
      if (rc || used) return rc;
    }
    return mnf_tc_paired_forward(ctx, 1, a.M, a.K, a.N, x, z, mean_W, log_std_W, mean_b, log_var_b, max_std, a.eps, y,
                                 mean_out, sd_out, geo);
  }
  rc = need_var();
  if (rc) return rc;
  return launch_paired<true>(ctx, a);
}

int mnf_conv2d_forward_mc(mnf_ctx* ctx, const mnf_conv_shape* s, const mnf_mc_source* src, const float* z, const float* mean_W,
                          const float* log_std_W, const float* mean_b, const float* log_var_b, float max_std,
                          const mnf_noise* eps, float* y, int32_t* used) {
  MNF_CHECK_ARG(ctx && s && src && mean_W && log_std_W && mean_b && log_var_b && y && eps && used, "mnf_conv2d_forward_mc: NULL argument");
  MNF_CHECK_ARG(src->mean && src->sd && src->n_rep >= 1 && src->Ho >= 2 && src->Wo >= 2 && src->F >= 1, "mnf_conv2d_forward_mc: bad source");
  *used = 0;
  int64_t Ho, Wo;
  int rc = mnf_conv_out_hw(s, &Ho, &Wo);
  if (rc) return rc;
  if (s->B == 0 || ctx->math_mode == MNF_MATH_FP32 || s->same || !s->fuse_relu_pool || s->H * 2 != src->Ho || s->W * 2 != src->Wo ||
      s->C != src->F || !mnf_tc_k_ok(ctx, s->kh * s->kw * s->C) || getenv("MNF_NO_F16") || getenv("MNF_NO_MC_FUSE"))
    return MNF_OK;
  int32_t ok = 0;
  rc = mnf_conv2d_can_fuse_pool(ctx, s, &ok);
  if (rc || !ok) return rc;
  const int geo[11] = {(int)s->H, (int)s->W, (int)s->C, (int)s->kh, (int)s->kw, (int)Ho, (int)Wo, (int)s->stride, 0, 0, 1};
  bool u = false;
  rc = mnf_conv_raw_f16_forward(ctx, s->B, geo, (int)s->F, nullptr, z, mean_W, log_std_W, mean_b, log_var_b, max_std, make_noise(eps), y,
                                nullptr, nullptr, &u, src);
  *used = u ? 1 : 0;
  return rc;
}

}  // extern "C"
