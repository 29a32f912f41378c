// Row-tile kernel for one IAF / RNVP flow step (included by flows.cu after FlowStepArgs).
// A CTA owns 64 rows.  Both MLP layers run as register-blocked FFMA GEMMs out of shared
// memory (4x4 resp. 4x(2+2) outputs per thread, operands fetched with 128-bit LDS), the
// hidden activations stay in shared memory, and the per-row log-det is reduced with warp
// shuffles.  z is read twice from global (stage 1 and the combine), the new state written once.
#pragma once

#define FT_ROWS 64
#define FT_THREADS 256
#define FT_DC 32          // reduction / column chunk
#define FT_HMAX 64        // hidden units supported (padded to 64 in shared memory)
#define FT_LD (FT_ROWS + 4)

template <int KIND>
__global__ void __launch_bounds__(FT_THREADS) flow_mlp_step_fwd(FlowStepArgs a) {
  // stage-1 operands and stage-2 weight chunks share one buffer
  __shared__ __align__(16) float ubuf[2][FT_DC][FT_LD];   // [0]: z chunk [dd][row] | Ws chunk [h][c..]
  __shared__ __align__(16) float hs[FT_HMAX][FT_LD];      // hidden activations, k-major: [h][row]
  float (*zs)[FT_LD] = ubuf[0];
  float (*w1s)[FT_LD] = ubuf[1];                           // [dd][h]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 thread grid
  const int64_t row0 = (int64_t)blockIdx.x * FT_ROWS;
  const int D = a.D, H = a.H;

  // ---- stage 1: h = act(in @ W1 + b1)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int d0 = 0; d0 < D; d0 += FT_DC) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int e = tid + FT_THREADS * q;
      int dd = e & 31, rr = e >> 5;
      int64_t grow = row0 + rr;
      int d = d0 + dd;
      float v = 0.f;
      if (grow < a.B && d < D) {
        v = load_zin(a, grow, d);
        if (a.z0_out && !a.zin) a.z0_out[grow * D + d] = v;
        if (KIND == MNF_FLOW_RNVP) v *= noise_bern1(a.bern, grow, (uint32_t)d, D, 0.5f);
      }
      zs[dd][rr] = v;
    }
    for (int e = tid; e < FT_DC * FT_HMAX; e += FT_THREADS) {
      int dd = e >> 6, h = e & 63;
      int d = d0 + dd;
      float w = 0.f;
      if (d < D && h < H) {
        w = __ldg(a.w1 + (int64_t)d * H + h);
        if (a.m1) w *= __ldg(a.m1 + (int64_t)d * H + h);
      }
      w1s[dd][h] = w;
    }
    __syncthreads();
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
  __syncthreads();

  // ---- stage 2: (s,t) = h @ [Ws|Wt] + [bs|bt]; thread = 4 rows x 2 columns x (s,t)
  // the stage-1 buffer is re-carved as two [FT_HMAX][34] weight chunks (64*34*2 == 2*32*68 floats)
  constexpr int WLD = FT_DC + 2;
  static_assert(2 * FT_HMAX * WLD <= 2 * FT_DC * FT_LD, "stage-2 chunks must fit the stage-1 buffer");
  float* wss = &ubuf[0][0][0];
  float* wts = wss + FT_HMAX * WLD;
  float ldrow[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < D; c0 += FT_DC) {
    for (int e = tid; e < FT_HMAX * FT_DC; e += FT_THREADS) {
      int h = e >> 5, cc = e & 31;
      int c = c0 + cc;
      float vs = 0.f, vt = 0.f;
      if (c < D && h < H) {
        vs = __ldg(a.ws + (int64_t)h * a.ld2 + c);
        vt = __ldg(a.wt + (int64_t)h * a.ld2 + c);
        if (a.m2) {
          vs *= __ldg(a.m2 + (int64_t)h * a.ld2 + c);
          vt *= __ldg(a.m2 + (int64_t)h * a.ld2 + D + c);
        }
      }
      wss[h * WLD + cc] = vs;
      wts[h * WLD + cc] = vt;
    }
    __syncthreads();
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
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int c = c0 + tx * 2 + j;
      if (c >= D) continue;
      float bs = a.bs[c], bt = a.bt[c];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int64_t row = row0 + ty * 4 + i;
        if (row >= a.B) continue;
        float sv = s[i][j] + bs, tv = t[i][j] + bt;
        float zv = a.zin ? a.zin[row * D + c] : load_zin(a, row, c);
        if (KIND == MNF_FLOW_IAF) {
          a.zout[row * D + (a.parity ? D - 1 - c : c)] = fmaf(zv, expf(sv), tv);
          ldrow[i] += sv;
        } else {
          float m = noise_bern1(a.bern, row, (uint32_t)c, D, 0.5f);
          float gate = 1.f / (1.f + expf(-sv));
          float lg = sv >= 0.f ? -log1pf(expf(-sv)) : sv - log1pf(expf(sv));  // log(sigmoid(sv))
          a.zout[row * D + c] = ((1.f - m) * zv * gate + (1.f - gate) * tv) + m * zv;
          ldrow[i] += (1.f - m) * lg;
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
