// Row-tile kernel for one IAF / RNVP flow step (included by flows.cu after FlowStepArgs).
// A CTA owns 32 rows (128 threads).  Both MLP layers run as register-blocked FFMA GEMMs out of
// shared memory (4x4 resp. 4x(2+2) outputs per thread, operands fetched with 128/64-bit LDS);
// the hidden activations stay in shared memory; the per-row log-det is reduced with warp
// shuffles.  Global loads are issued in batches into registers one chunk AHEAD of the math
// (software pipeline), and one Philox call serves four Bernoulli / normal draws.
#pragma once

#define FT_ROWS 32
#define FT_THREADS 128
#define FT_DC 32          // reduction / column chunk
#define FT_HMAX 64        // hidden units supported (padded to 64 in shared memory)
#define FT_LD (FT_ROWS + 4)

// four Bernoulli(p) draws for elements [4*g4, 4*g4+4) of row lrow
__device__ __forceinline__ void noise_bern4(const NoiseDev& n, int64_t lrow, uint32_t g4, int n_inner, float p,
                                            float out[4]) {
  if (n.injected) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int d = (int)(g4 * 4 + e);
      out[e] = d < n_inner ? n.injected[lrow * n_inner + d] : 0.f;
    }
    return;
  }
  if (p == 0.5f) {  // fair coins: one bit per element (common.cuh)
    const Philox4 c = philox_at(n, n.row_offset + (uint64_t)lrow, 0u, g4 >> 5);
#pragma unroll
    for (int e = 0; e < 4; ++e) out[e] = coin_of(c, g4 * 4 + e);
    return;
  }
  Philox4 q = philox_at(n, n.row_offset + (uint64_t)lrow, 0u, g4);
  out[0] = u01_open(q.x) <= p ? 1.f : 0.f;
  out[1] = u01_open(q.y) <= p ? 1.f : 0.f;
  out[2] = u01_open(q.z) <= p ? 1.f : 0.f;
  out[3] = u01_open(q.w) <= p ? 1.f : 0.f;
}

// z (or the sampled z0) for elements [4*g4, 4*g4+4) of one row
__device__ __forceinline__ void load_z4(const FlowStepArgs& a, int64_t row, int g4, float out[4]) {
  const int D = a.D, d = g4 * 4;
  if (a.zin) {
    const float* p = a.zin + row * D + d;
    if (d + 3 < D && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      float4 v = __ldg(reinterpret_cast<const float4*>(p));
      out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) out[e] = d + e < D ? __ldg(p + e) : 0.f;
    }
    return;
  }
  float nrm[4];
  if (a.eps.injected) {
#pragma unroll
    for (int e = 0; e < 4; ++e) nrm[e] = d + e < D ? a.eps.injected[row * D + d + e] : 0.f;
  } else {
    philox_normal4(a.eps, a.eps.row_offset + (uint64_t)row, 0u, (uint32_t)g4, nrm);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
    out[e] = d + e < D ? a.q0_mean[d + e] + expf(0.5f * a.q0_log_var[d + e]) * nrm[e] : 0.f;
}

template <int KIND>
__global__ void __launch_bounds__(FT_THREADS) flow_mlp_step_fwd(FlowStepArgs a) {
  constexpr int WLD = FT_DC + 2;        // stage-2 weight chunk row stride
  constexpr int W1LD = FT_HMAX + 4;     // stage-1 weight chunk row stride
  constexpr int S1 = FT_DC * FT_LD + FT_DC * W1LD, S2 = 2 * FT_HMAX * WLD;
  __shared__ __align__(16) float ubuf[S1 > S2 ? S1 : S2];
  __shared__ __align__(16) float hs[FT_HMAX][FT_LD];      // hidden activations, k-major: [h][row]
  float (*zs)[FT_LD] = reinterpret_cast<float (*)[FT_LD]>(ubuf);                    // [dd][row]
  float (*w1s)[W1LD] = reinterpret_cast<float (*)[W1LD]>(ubuf + FT_DC * FT_LD);    // [dd][h]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 8 thread grid
  const int64_t row0 = (int64_t)blockIdx.x * FT_ROWS;
  const int D = a.D, H = a.H;

  // ---- stage 1: h = act(in @ W1 + b1).  Loader: thread -> 2 x (row, 4 consecutive d)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float zreg[2][4], wreg[16];
  auto s1_load = [&](int d0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = tid + FT_THREADS * q, rr = e >> 3, dq = e & 7;
      const int64_t grow = row0 + rr;
      const int g4 = (d0 >> 2) + dq;
      if (grow < a.B && g4 * 4 < D) {
        load_z4(a, grow, g4, zreg[q]);
        if (a.z0_out && !a.zin) {
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2)
            if (g4 * 4 + e2 < D) a.z0_out[grow * D + g4 * 4 + e2] = zreg[q][e2];
        }
        if (KIND == MNF_FLOW_RNVP) {
          float m[4];
          noise_bern4(a.bern, grow, (uint32_t)g4, D, 0.5f, m);
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) zreg[q][e2] *= m[e2];
        }
      } else {
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) zreg[q][e2] = 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = tid + FT_THREADS * q, dd = e >> 6, h = e & 63, d = d0 + dd;
      float w = 0.f;
      if (d < D && h < H) {
        w = __ldg(a.w1 + (int64_t)d * H + h);
        if (a.m1) w *= __ldg(a.m1 + (int64_t)d * H + h);
      }
      wreg[q] = w;
    }
  };
  auto s1_store = [&]() {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = tid + FT_THREADS * q, rr = e >> 3, dq = e & 7;
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) zs[dq * 4 + e2][rr] = zreg[q][e2];
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = tid + FT_THREADS * q;
      w1s[e >> 6][e & 63] = wreg[q];
    }
  };
  s1_load(0);
  for (int d0 = 0; d0 < D; d0 += FT_DC) {
    s1_store();
    __syncthreads();
    if (d0 + FT_DC < D) s1_load(d0 + FT_DC);
#pragma unroll 8
    for (int dd = 0; dd < FT_DC; ++dd) {
      float4 zv = *reinterpret_cast<const float4*>(&zs[dd][ty * 4]);
      float4 wv = *reinterpret_cast<const float4*>(&w1s[dd][tx * 4]);
      float zz[4] = {zv.x, zv.y, zv.z, zv.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(zz[i], ww[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int h = tx * 4 + j;
    float b = h < H ? a.b1[h] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int rr = ty * 4 + i;
      float pre = acc[i][j] + b;
      float hv = h < H ? ((KIND == MNF_FLOW_IAF) ? fmaxf(pre, 0.f) : tanhf(pre)) : 0.f;
      hs[h][rr] = hv;
      if (a.hid_save && h < H && row0 + rr < a.B) a.hid_save[(row0 + rr) * H + h] = hv;
    }
  }

  // ---- stage 2: (s,t) = h @ [Ws|Wt] + [bs|bt]; thread = 4 rows x 2 columns x (s,t)
  // the stage-1 buffer is re-carved as two [FT_HMAX][34] weight chunks
  float* wss = ubuf;
  float* wts = wss + FT_HMAX * WLD;
  float ldrow[4] = {0.f, 0.f, 0.f, 0.f};
  float sreg[16], treg[16], zin2[4][2];
  auto s2_load = [&](int c0) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = tid + FT_THREADS * q, h = e >> 5, c = c0 + (e & 31);
      float vs = 0.f, vt = 0.f;
      if (c < D && h < H) {
        vs = __ldg(a.ws + (int64_t)h * a.ld2 + c);
        vt = __ldg(a.wt + (int64_t)h * a.ld2 + c);
        if (a.m2) {
          vs *= __ldg(a.m2 + (int64_t)h * a.ld2 + c);
          vt *= __ldg(a.m2 + (int64_t)h * a.ld2 + D + c);
        }
      }
      sreg[q] = vs;
      treg[q] = vt;
    }
  };
  auto s2_store = [&]() {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = tid + FT_THREADS * q;
      wss[(e >> 5) * WLD + (e & 31)] = sreg[q];
      wts[(e >> 5) * WLD + (e & 31)] = treg[q];
    }
  };
  // this thread's z inputs for the combine step of chunk c0 (recomputed when z0 is sampled)
  auto z_load = [&](int c0) {
    const int c = c0 + tx * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t row = row0 + ty * 4 + i;
      zin2[i][0] = zin2[i][1] = 0.f;
      if (row < a.B && c < D) {
        if (a.zin) {
          zin2[i][0] = __ldg(a.zin + row * D + c);
          if (c + 1 < D) zin2[i][1] = __ldg(a.zin + row * D + c + 1);
        } else {
          float z4[4];
          load_z4(a, row, c >> 2, z4);
          zin2[i][0] = z4[c & 3];
          zin2[i][1] = z4[(c & 3) + 1];  // c is even: c & 3 is 0 or 2
        }
      }
    }
  };
  s2_load(0);
  __syncthreads();  // hs complete, stage-1 buffers free
  for (int c0 = 0; c0 < D; c0 += FT_DC) {
    s2_store();
    __syncthreads();
    z_load(c0);
    if (c0 + FT_DC < D) s2_load(c0 + FT_DC);
    float s[4][2], t[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = t[i][0] = t[i][1] = 0.f;
#pragma unroll 10
    for (int h = 0; h < H; ++h) {
      float4 hv = *reinterpret_cast<const float4*>(&hs[h][ty * 4]);
      float2 sv = *reinterpret_cast<const float2*>(&wss[h * WLD + tx * 2]);
      float2 tv = *reinterpret_cast<const float2*>(&wts[h * WLD + tx * 2]);
      float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i][0] = fmaf(hh[i], sv.x, s[i][0]);
        s[i][1] = fmaf(hh[i], sv.y, s[i][1]);
        t[i][0] = fmaf(hh[i], tv.x, t[i][0]);
        t[i][1] = fmaf(hh[i], tv.y, t[i][1]);
      }
    }
    const int cb = c0 + tx * 2;
    if (cb < D) {
      float bsv[2], btv[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        bsv[j] = cb + j < D ? a.bs[cb + j] : 0.f;
        btv[j] = cb + j < D ? a.bt[cb + j] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t row = row0 + ty * 4 + i;
        if (row >= a.B) continue;
        float m4[4];
        if (KIND == MNF_FLOW_RNVP) noise_bern4(a.bern, row, (uint32_t)(cb >> 2), D, 0.5f, m4);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = cb + j;
          if (c >= D) continue;
          const float sv = s[i][j] + bsv[j], tv = t[i][j] + btv[j], zv = zin2[i][j];
          if (KIND == MNF_FLOW_IAF) {
            a.zout[row * D + (a.parity ? D - 1 - c : c)] = fmaf(zv, expf(sv), tv);
            ldrow[i] += sv;
          } else {
            const float m = m4[(cb & 3) + j];
            const float gate = 1.f / (1.f + expf(-sv));
            const float lg = sv >= 0.f ? -log1pf(expf(-sv)) : sv - log1pf(expf(sv));  // log(sigmoid(sv))
            a.zout[row * D + c] = ((1.f - m) * zv * gate + (1.f - gate) * tv) + m * zv;
            ldrow[i] += (1.f - m) * lg;
          }
        }
      }
    }
    __syncthreads();
  }
  // the 16 threads sharing a row group are the 16 lanes of a half-warp
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = ldrow[i];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    int64_t row = row0 + ty * 4 + i;
    if (tx == 0 && row < a.B && a.ld_out) a.ld_out[row] = (a.ld_in ? a.ld_in[row] : 0.f) + v;
  }
}
