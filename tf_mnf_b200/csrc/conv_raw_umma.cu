// MNFConv2D.call (mnf_conv.py:182-197) on tcgen05 WITHOUT materialising im2col.
//
// For stride-1 convolutions whose output rows are 8 pixels wide (MNF-LeNet conv2: 12x12x20 ->
// 8x8x50) the A operand of every filter tap is a *strided view* of the raw input tile, and the
// UMMA shared-memory descriptor can express that view directly:
//
//   planes[v][cc][hi][img][wi] of 16-byte chunks (4 channels), v in {x hi, x lo, x^2 hi, x^2 lo}
//   tile rows r = (ho*I + img)*8 + wo   (I images per 128-row tile)
//   tap (i,j), channel chunk cc:  row r lives at  cc*CCS + ((ho+i)*I + img)*Wi*16 + (wo+j)*16
//     -> the 8 rows of a core matrix (wo = 0..7) are 8 consecutive 16-byte chunks   (K-major,
//        SWIZZLE_NONE), consecutive core matrices are Wi*16 bytes apart (SBO), and the two
//        16-byte K chunks of one K=8 MMA are `lbo` bytes apart -- any positive multiple of 16,
//        chosen PER MMA from a host-built step table (two chunks of one tap: lbo = CCS; the odd
//        leftover chunks of two neighbouring taps: lbo = their pixel distance).
//
// So every input element is split (x -> hi/lo, x^2 -> hi/lo; 3xTF32) exactly ONCE per tile instead
// of once per filter tap (25x for a 5x5 filter), and the producer warps do ~20x less work than
// the gather in tc_gemm.cu.  The mean and the variance products run as two phases per tile so
// the planes of one phase are refilled while the tensor core works on the other.
//
//   warps 0-3, 8-11  epilogue, columns [0,32) resp. [32,64) of a row (TMEM -> bias, z-scale,
//                    Philox eps, optional ReLU+2x2 max-pool, staged coalesced store)
//   warp  4          MMA issue (3 tcgen05.mma.kind::tf32 per step and phase), tcgen05.commit
//   warps 5-7        plane producers;  warp 12 lane 0: weight-block loader (cp.async.bulk ring)
#include <utility>
#include <vector>

#include "umma.cuh"

namespace cr {
using namespace um;

constexpr int BM = 128, BN = 64;
constexpr int G = 6;                 // k8-steps per weight stage
constexpr int STEP_B = 2 * 2 * BN * 16;  // bytes of one step's weights for one phase: [2 kc][hi 64 n | lo 64 n][16 B] = 4 KB
constexpr int B_STAGE = G * STEP_B;  // 24 KB
constexpr int SB = 3;                // weight ring depth
constexpr int THREADS = 13 * 32;
constexpr int PROD_WARPS = 3;      // warps 5-7 fill the planes; warps 8-11 are the second epilogue group
constexpr int TMEM_COLS = 512;  // [2 buffers][2 phases][hi-product 64 | lo-product 64]

struct Args {
  int64_t B;         // images
  int H, W, C, kh, kw, Ho, N, I;  // Wo == 8; I images per tile
  int Hi, Wi;        // plane extent per image: Ho-1+kh, 8-1+kw
  int cpt;           // 16-byte channel chunks per pixel (C / 4)
  int steps;         // k8 steps
  int pool;
  const float* x;
  const float* z;    // [B,N] or nullptr
  const float* bm;
  const float* vb;
  const int* steptab;  // [steps][2]: a_off, a_lbo (bytes)
  const float* Bp;     // [2 phases][steps][STEP_B/4 floats]
  float* y;
  float* mean_out;
  float* sd_out;
  NoiseDev eps;
};

__global__ void __launch_bounds__(THREADS, 1) conv_raw_umma(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int pix = a.Hi * a.I * a.Wi;          // 16-byte chunks per channel-chunk plane
  const int CCS = pix * 16;                   // bytes between channel chunks
  const int PLANE = a.cpt * CCS;              // bytes of one variant
  uint8_t* planes = smem;                     // [4][cpt][Hi][I][Wi][16 B]
  uint8_t* Bs = planes + 4 * PLANE;           // [SB][B_STAGE]
  float* ys = (float*)(Bs + SB * B_STAGE);    // [BM][N] dense staging
  int* stab = (int*)(ys + BM * BN);           // [steps][2]
  float* bias_s = (float*)(stab + 2 * a.steps);  // [2][BN]
  float* zs = bias_s + 2 * BN;                   // [I][BN] z rows of the tile being written out
  uint64_t* bars = (uint64_t*)(((uintptr_t)(zs + 16 * BN) + 7) & ~(uintptr_t)7);
  uint64_t* pfull = bars;        // [2] planes of phase ready
  uint64_t* pempty = pfull + 2;  // [2]
  uint64_t* bfull = pempty + 2;  // [SB]
  uint64_t* bempty = bfull + SB; // [SB]
  uint64_t* tfull = bempty + SB; // [2]
  uint64_t* tempty = tfull + 2;  // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HoWo = a.Ho * 8;
  const int64_t ntiles = (a.B + a.I - 1) / a.I;
  const int nstages = (a.steps + G - 1) / G;

  if (threadIdx.x == 0) {
    for (int p = 0; p < 2; ++p) { mbar_init(&pfull[p], PROD_WARPS); mbar_init(&pempty[p], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * a.steps; i += THREADS) stab[i] = a.steptab[i];
  for (int n = threadIdx.x; n < BN; n += THREADS) {
    bias_s[n] = n < a.N ? a.bm[n] : 0.f;
    bias_s[BN + n] = n < a.N ? a.vb[n] : 1.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4 || (warp >= 8 && warp < 12)) {
    // ================================================================== epilogue warps
    const int quad = warp & 3, chalf = warp >= 8 ? 1 : 0;
    const int et = chalf * 128 + quad * 32 + lane;   // 0..255 among the epilogue threads
    const int r = quad * 32 + lane;             // TMEM lane = tile row (ho, img, wo)
    const int wo = r & 7, g = r >> 3, img = g % a.I, ho = g / a.I;
    const int rl = img * HoWo + ho * 8 + wo;    // row inside the tile's block of y
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int64_t b0 = t * a.I, brow = b0 + img;
      const int buf = it & 1;
      // this tile's z rows -> shared memory (the previous tile's readers passed the closing barrier)
      for (int e = et; e < a.I * BN; e += 256) {
        const int im = e / BN, n = e - im * BN;
        zs[e] = (a.z && b0 + im < a.B && n < a.N) ? __ldg(a.z + (b0 + im) * a.N + n) : 1.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull[buf], (it >> 1) & 1);
      tc_fence_after();
      const bool row_ok = brow < a.B;
      const uint32_t outer = (uint32_t)(ho * 8 + wo);
      const int64_t m = brow * HoWo + outer;
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 4 * BN);
#pragma unroll 1
      for (int c0 = chalf * 32; c0 < chalf * 32 + 32; c0 += 16) {
        if (c0 >= a.N) break;
        float mv[16], vv[16];
        {
          float m2[16], v2[16];
          tmem_ld16(trow + c0, mv);
          tmem_ld16(trow + BN + c0, m2);
          tmem_ld16(trow + 2 * BN + c0, vv);
          tmem_ld16(trow + 3 * BN + c0, v2);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            mv[j] += m2[j];
            vv[j] += v2[j];
          }
        }
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int nb = c0 + 4 * q;
            if (nb >= a.N) break;
            float e[4];
            if (a.eps.injected) {
#pragma unroll
              for (int j = 0; j < 4; ++j) e[j] = nb + j < a.N ? a.eps.injected[m * a.N + nb + j] : 0.f;
            } else {
              philox_normal4(a.eps, a.eps.row_offset + (uint64_t)brow, outer, (uint32_t)(nb >> 2), e);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = nb + j;
              if (n < a.N) {
                float mean = mv[4 * q + j] + bias_s[n];
                const float sd = sqrtf(vv[4 * q + j] + bias_s[BN + n]);
                if (a.sd_out) a.sd_out[m * a.N + n] = sd;
                if (a.mean_out) a.mean_out[m * a.N + n] = mean;
                if (a.z) mean *= zs[img * BN + n];
                ys[rl * a.N + n] = fmaf(sd, e[j], mean);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int64_t nimg = (a.B - b0 < a.I) ? a.B - b0 : a.I;
      const int rows = (int)nimg * HoWo;
      if (a.pool) {
        const int np = rows >> 2;
        float* dst = a.y + (b0 * HoWo >> 2) * a.N;
        for (int e = et; e < np * a.N; e += 256) {
          const int pp = e / a.N, c = e - pp * a.N;
          const int hp = pp >> 2, wp = pp & 3;  // Wo / 2 == 4 pooled columns
          const float* s0 = ys + ((2 * hp) * 8 + 2 * wp) * a.N + c;
          const float v = fmaxf(fmaxf(s0[0], s0[a.N]), fmaxf(s0[8 * a.N], s0[9 * a.N]));
          dst[e] = fmaxf(v, 0.f);
        }
      } else {
        float* dst = a.y + b0 * HoWo * a.N;
        for (int e = et; e < rows * a.N; e += 256) dst[e] = ys[e];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  } else if (warp == 4) {
    // ======================================================================= MMA issuer
    // x_hi * [W_hi | W_lo] is ONE N=128 instruction (the A tile is read from shared memory once for
    // both products: at N=64 the operand fetch, not the tensor pipe, sets the pace); x_lo * W_hi
    // is an N=64 instruction into the first 64 columns.  The epilogue adds the two column halves.
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t sbo = (uint32_t)(a.Wi * 16);
    const uint64_t bdesc0 = make_desc(smem_u32(Bs), 2 * BN * 16, 128);
    const uint32_t plane0 = smem_u32(planes);
    int it = 0, st = 0;
    uint32_t bph = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
      for (int ph = 0; ph < 2; ++ph) {
        mbar_wait(&pfull[ph], it & 1);
        tc_fence_after();
        const uint32_t dacc = tmem_base + (uint32_t)(buf * 4 * BN + ph * 2 * BN);
        const uint32_t pa_hi = plane0 + (uint32_t)((2 * ph) * PLANE), pa_lo = pa_hi + (uint32_t)PLANE;
        for (int sg = 0; sg < nstages; ++sg) {
            mbar_wait(&bfull[st], bph);
          tc_fence_after();
          if (lane == 0) {
            const int s1 = min(a.steps, (sg + 1) * G);
            for (int s = sg * G; s < s1; ++s) {
              const uint32_t aoff = (uint32_t)stab[2 * s], lbo = (uint32_t)stab[2 * s + 1];
              const uint64_t ah = make_desc(pa_hi + aoff, lbo, sbo), al = make_desc(pa_lo + aoff, lbo, sbo);
              const uint64_t bh = bdesc0 + (uint64_t)((st * B_STAGE + (s - sg * G) * STEP_B) >> 4);
              umma_tf32(dacc, ah, bh, idesc2, s == 0 ? 0u : 1u);
              umma_tf32(dacc, al, bh, idesc, 1u);
            }
            umma_commit(&bempty[st]);
            if (sg == nstages - 1) {
              umma_commit(&pempty[ph]);              // planes of this phase may be refilled
              if (ph == 1) umma_commit(&tfull[buf]);  // both accumulators complete
            }
          }
          __syncwarp();
          if (++st == SB) { st = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 12) {
    // ==================================================================== weight loader
    if (lane == 0) {
      int st = 0;
      uint32_t bph = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int ph = 0; ph < 2; ++ph) {
          const float* src0 = a.Bp + (size_t)ph * a.steps * (STEP_B / 4);
          for (int sg = 0; sg < nstages; ++sg) {
            const int ns = min(G, a.steps - sg * G);
            mbar_wait(&bempty[st], bph ^ 1);
            mbar_arrive_expect_tx(&bfull[st], (uint32_t)(ns * STEP_B));
            bulk_g2s(Bs + st * B_STAGE, src0 + (size_t)sg * G * (STEP_B / 4), (uint32_t)(ns * STEP_B), &bfull[st]);
            if (++st == SB) { st = 0; bph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================== plane producers
    if (warp < 8) {
      const int pt = threadIdx.x - 5 * 32;  // 0 .. 95
      constexpr int NP = PROD_WARPS * 32;
      const int nchunks = a.cpt * pix;       // 16-byte chunks per variant
      const int img_f = a.H * a.W * a.C;     // floats per input image
      int it = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int64_t b0 = t * a.I;
        for (int ph = 0; ph < 2; ++ph) {
          mbar_wait(&pempty[ph], (it & 1) ^ 1);
          uint8_t* dst_hi = planes + (2 * ph) * PLANE;
          uint8_t* dst_lo = dst_hi + PLANE;
          for (int q0 = pt; q0 < nchunks; q0 += 4 * NP) {
            // four independent 16-byte loads in flight per thread, then split + store
            float4 v[4];
            int o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int q = q0 + u * NP;
              v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              o[u] = -1;
              if (q < nchunks) {
                // q enumerates (hi, img, wi, cc), cc fastest: a warp reads consecutive floats
                const int cc = q % a.cpt;
                int p = q / a.cpt;
                const int wi = p % a.Wi;
                p /= a.Wi;
                const int im = p % a.I, hi = p / a.I;
                int64_t b = b0 + im;
                if (b >= a.B) b = a.B - 1;
                if (hi < a.H && wi < a.W)
                  v[u] = __ldg(reinterpret_cast<const float4*>(a.x + b * img_f + (hi * a.W + wi) * a.C + cc * 4));
                o[u] = cc * CCS + ((hi * a.I + im) * a.Wi + wi) * 16;
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (o[u] < 0) continue;
              const float in[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
              float h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split_tf32(ph == 0 ? in[e] : in[e] * in[e], h[e], l[e]);
              *reinterpret_cast<float4*>(dst_hi + o[u]) = make_float4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<float4*>(dst_lo + o[u]) = make_float4(l[0], l[1], l[2], l[3]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pfull[ph]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// weights -> Bp[phase][step][kc 2][hi n 64 | lo n 64][4 k]; kmap[step][8] = global k index or -1
__global__ void pack_raw_weights(int N, int steps, const int* __restrict__ kmap, const float* __restrict__ Wm,
                                 const float* __restrict__ Ls, float max_std, float* __restrict__ Bp,
                                 const float* __restrict__ lvb, float* __restrict__ vb) {
  const int64_t total = (int64_t)2 * steps * 2 * BN * 4;  // (phase, step, kc, n, e)
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx & 3);
    int64_t t = idx >> 2;
    const int n = (int)(t % BN);
    t /= BN;
    const int kc = (int)(t & 1);
    t >>= 1;
    const int s = (int)(t % steps), ph = (int)(t / steps);
    const int k = kmap[s * 8 + kc * 4 + e];
    float w = 0.f;
    if (k >= 0 && n < N) {
      if (ph == 0) w = Wm[(int64_t)k * N + n];
      else {
        const float sd = fminf(expf(Ls[(int64_t)k * N + n]), max_std);
        w = sd * sd;
      }
    }
    float hi, lo;
    split_tf32(w, hi, lo);
    float* base = Bp + ((int64_t)ph * steps + s) * (STEP_B / 4) + ((int64_t)kc * 2 * BN + n) * 4 + e;
    base[0] = hi;
    base[BN * 4] = lo;  // the lo rows follow the 64 hi rows of the same K chunk
    if (idx < N) vb[idx] = fminf(expf(lvb[idx]), max_std * max_std);
  }
}

}  // namespace cr

// Returns MNF_OK with *used = false when the shape is outside this kernel's domain.
int mnf_conv_raw_umma_forward(mnf_ctx* ctx, int64_t B, const int* geo, int N, const float* x, const float* z,
                              const float* mean_W, const float* log_std_W, const float* mean_b,
                              const float* log_var_b, float max_std, const NoiseDev& eps, float* y, float* mean_out,
                              float* sd_out, bool* used) {
  using namespace cr;
  *used = false;
  const int H = geo[0], W = geo[1], C = geo[2], kh = geo[3], kw = geo[4], Ho = geo[5], Wo = geo[6], stride = geo[7],
            pad_t = geo[8], pad_l = geo[9], pool = geo[10];
  if (stride != 1 || Wo != 8 || pad_t || pad_l || (C & 3) || N > BN || N <= 20) return MNF_OK;
  if (Ho < 1 || Ho > 16 || (16 % Ho) != 0) return MNF_OK;
  if (((uintptr_t)x & 15) != 0) return MNF_OK;
  Args a = {};
  a.B = B; a.H = H; a.W = W; a.C = C; a.kh = kh; a.kw = kw; a.Ho = Ho; a.N = N;
  a.I = 16 / Ho;
  a.Hi = Ho - 1 + kh; a.Wi = 8 - 1 + kw;
  if (a.Hi > H || a.Wi > W) return MNF_OK;  // VALID geometry only
  a.cpt = C / 4;
  a.pool = pool;
  // ---- step table: chunk pairs of one tap, then the odd leftover chunks of neighbouring taps
  const int pixn = a.Hi * a.I * a.Wi, CCS = pixn * 16;
  std::vector<int> tab, kmap;
  auto chunk_off = [&](int i, int j, int cc) { return cc * CCS + (i * a.I * a.Wi + j) * 16; };
  auto push_k = [&](int i, int j, int cc) {
    for (int e = 0; e < 4; ++e) kmap.push_back(i >= 0 ? (i * kw + j) * C + cc * 4 + e : -1);
  };
  std::vector<std::pair<int, int>> left;
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j) {
      for (int cc = 0; cc + 1 < a.cpt; cc += 2) {
        tab.push_back(chunk_off(i, j, cc));
        tab.push_back(CCS);
        push_k(i, j, cc);
        push_k(i, j, cc + 1);
      }
      if (a.cpt & 1) left.push_back({i, j});
    }
  for (size_t m = 0; m < left.size(); m += 2) {
    const int cc = a.cpt - 1;
    const int o0 = chunk_off(left[m].first, left[m].second, cc);
    tab.push_back(o0);
    push_k(left[m].first, left[m].second, cc);
    if (m + 1 < left.size()) {
      const int o1 = chunk_off(left[m + 1].first, left[m + 1].second, cc);
      if (o1 <= o0 || (o1 - o0) >= (1 << 18)) return MNF_OK;
      tab.push_back(o1 - o0);
      push_k(left[m + 1].first, left[m + 1].second, cc);
    } else {
      tab.push_back(16);   // dummy second chunk: valid memory, zero weights
      push_k(-1, 0, 0);
    }
  }
  a.steps = (int)tab.size() / 2;
  const size_t plane_bytes = (size_t)4 * a.cpt * CCS;
  const size_t smem = plane_bytes + (size_t)SB * B_STAGE + sizeof(float) * BM * BN + sizeof(int) * 2 * a.steps +
                      sizeof(float) * (2 + 16) * BN + 256;
  if (smem > 227 * 1024) return MNF_OK;
  // ---- workspace: Bp | vb | steptab | kmap
  const size_t bp_floats = (size_t)2 * a.steps * (STEP_B / 4);
  const size_t ws_bytes = sizeof(float) * (bp_floats + N + 64) + sizeof(int) * (tab.size() + kmap.size() + 64);
  char* ws = (char*)mnf_workspace(ctx, ws_bytes);
  if (!ws) return MNF_ERR_CUDA;
  float* Bp = (float*)ws;
  float* vb = Bp + bp_floats;
  int* d_tab = mnf_geom_table(ctx, 0xC0A32A, tab.data(), tab.size());  // shape-only tables: uploaded once per context
  int* d_kmap = mnf_geom_table(ctx, 0xC0A32B, kmap.data(), kmap.size());
  if (!d_tab || !d_kmap) return MNF_ERR_CUDA;
  const int64_t tot = (int64_t)2 * a.steps * 2 * BN;
  int grid = (int)(cdiv(tot, 256) < ctx->sm_count * 4 ? cdiv(tot, 256) : ctx->sm_count * 4);
  pack_raw_weights<<<grid, 256, 0, ctx->stream>>>(N, a.steps, d_kmap, mean_W, log_std_W, max_std, Bp, log_var_b, vb);
  MNF_LAUNCHED(ctx);
  a.x = x; a.z = z; a.bm = mean_b; a.vb = vb; a.steptab = d_tab; a.Bp = Bp;
  a.y = y; a.mean_out = mean_out; a.sd_out = sd_out; a.eps = eps;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MNF_CUDA(cudaFuncSetAttribute(conv_raw_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int64_t tiles = cdiv(B, a.I);
  const int g = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
  const bool probed = mnf_probe_begin(ctx, MNF_PROBE_CONV_GEMM);
  mnf_note_path("conv_raw_umma");
  conv_raw_umma<<<g, THREADS, smem, ctx->stream>>>(a);
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  *used = true;
  return MNF_OK;
}
