// MNFDense.call (mnf_dense.py:159-170) for MANY rows under the fp32-parity tensor-core mode: the
// paired mean / variance contraction
//     mean = (x * z) W_mu            var = (x * x) min(exp(L_sigma), s_max)^2
// on tcgen05 kind::f16 with every operand split x = hi + lo into two fp16 planes (11 significant
// bits each; hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM: per-term error ~2^-22), both
// operand streams PACKED ONCE per call into the no-swizzle K-major UMMA layout, so the persistent GEMM
// kernel's producer is two cp.async.bulk per 32-wide k-stage and no thread of the GEMM touches an
// operand element.
//
// Why (measured on the tile kernel this replaces for M >= 1024, tc::paired_umma_fwd<128>): with
// producer warps converting fp32 x, z into split planes inside the GEMM, MNF-LeNet's d1 (10240 x 800
// x 500) redid that conversion once per 128-column tile (4 x), its MMA warp waited 52 % of the time
// for them, and 1.05 GB went from L2 to the SMs per launch for 56 MB of algorithmic bytes.  fp16
// halves the bytes per split element against TF32, the pack kernel reads x and z exactly once, and
// the GEMM streams 1.6 MB per 128 x 128 tile (0.52 GB per launch) with nothing but bulk copies.
//
// fp16 lacks range, so every operand gets an exact power-of-two scale (as conv_raw_f16.cu):
//   rows:    2^e with max_k |x z| * 2^e in [2^13, 2^14) and likewise for x^2  (per row, both kept)
//   columns: 2^e with max_k |W| * 2^e in [2^13, 2^14), mean and variance weights separately
// The epilogue multiplies the fp32 accumulators by the inverse scales (exact).  Elements far below
// their row / column maximum lose RELATIVE precision in the lo plane (fp16 subnormals) but their
// absolute error stays below 2^-38 of that maximum.
//
//   d3::row_pack      x, z -> Ap[m-tile][k-stage][{xz hi, xz lo, x^2 hi, x^2 lo}][4 chunks][128 rows][8 halves], rsc[m][2]
//   d3::col_scale / d3::weight_pack   W, L -> Bp[n-tile][k-stage][{W hi, W lo, V hi, V lo}][4][128][8], wsc[2][N], vb[N]
//   d3::paired_f16x3  warps 0-3 epilogue, warp 4 MMA issue, warp 5 bulk-copy producer; 3 x 64 KB stages;
//                     TMEM 512 columns = 2 buffers x {mean, var} x 128
#include <cuda_fp16.h>

#include "umma.cuh"

namespace d3 {
using namespace um;

constexpr int BM = 128, BN = 128;
constexpr int KS = 32;                      // k per pipeline stage
constexpr int CH = KS / 8;                  // 16-byte chunks (8 halves) per operand row and stage
constexpr int A_VAR = CH * BM * 16;         // 8 KB: one variant of one stage
constexpr int A_STAGE = 4 * A_VAR;          // 32 KB
constexpr int B_VAR = CH * BN * 16;
constexpr int B_STAGE = 4 * B_VAR;          // 32 KB
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 4;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int TMEM_COLS = 4 * BN;

struct Args {
  int64_t M, mtiles;
  int K, N, kblocks, ntiles;
  const uint8_t* Ap;
  const uint8_t* Bp;
  const float2* rsc;  // [mtiles * BM] inverse row scales {x z, x^2}
  const float* wsc;   // [2][ntiles * BN] inverse column scales {mean, var}
  const float* bm;
  const float* vb;
  float* y;
  float* sd_out;
  NoiseDev eps;
  int act;
  // K segments (contractions longer than MNF_TC_PARITY_KMAX in the fp32-parity mode): one launch covers k-blocks
  // [kb0, kb1) -- a TMEM accumulation chain of <= 1024 terms -- and the un-scaled partial sums travel between the
  // launches through `raw` ([2][M][N] fp32: mean, variance).  seg: 0 whole K in one launch, 1 first segment (raw =),
  // 2 middle (raw +=), 3 last (raw + this segment, then the epilogue)
  int kb0, kb1, seg;
  float* raw;
};

__device__ __forceinline__ void split_f16x2(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  // packed conversions (F2FP.F16.F32.PACK_AB: one instruction per PAIR, not on the quarter-rate XU pipe that the
  // scalar F2F.F16.F32 uses -- the plane producers convert 4 values per element)
  const __half2 h = __floats2half2_rn(a0, a1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 2^e with m * 2^e in [2^13, 2^14) (e = 0 for m == 0 / non-finite)
__device__ __forceinline__ int scale_exp(float m) {
  int e = 0;
  if (m > 0.f && m < 3.0e38f) {
    int ex;
    frexpf(m, &ex);  // m = f * 2^ex, f in [0.5, 1)
    e = 14 - ex;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
  }
  return e;
}

// ---- rows: one CTA = RP_ROWS rows of one 128-row tile, 8 warps x RP_ROWS / 8 rows.  The kernel is a chain of
// dependent load phases per CTA, so more, smaller CTAs in flight win: 41 us with 32 rows per CTA, 33 us with 16,
// 30 us with 8 (MNF-LeNet d1).  Pass 1: a warp per row finds
// max |x z| and max x^2 (coalesced float4 reads) -> the row's two scales.  Pass 2, per slab of 256 k:
// a lane converts ONE 16-byte chunk (8 consecutive k: 32 contiguous bytes of x and of z, the warp
// reads 1 KB contiguous; second visit, L1 / L2 hits) of all four variants and parks it in shared
// memory at [variant][chunk][row ^ chunk] (conflict-free both ways); then every (variant, chunk) line
// of 32 rows x 16 B leaves as ONE 512-byte coalesced store.  x and z are read from HBM once.
constexpr int RP_ROWS = 8, RP_THREADS = 256, RP_SLAB = 32;  // rows per CTA (8, 16 or 32), chunks per slab
constexpr int RP_PARK_BYTES = 4 * RP_SLAB * RP_ROWS * 16;
__global__ void __launch_bounds__(RP_THREADS) row_pack(int64_t M, int K, int kblocks, const float* __restrict__ x,
                                                       const float* __restrict__ z, uint8_t* __restrict__ Ap,
                                                       float2* __restrict__ rsc) {
  __shared__ float fsc[RP_ROWS][2];
  extern __shared__ __align__(16) uint8_t park_raw[];
  uint4 (*park)[RP_SLAB][RP_ROWS] = reinterpret_cast<uint4 (*)[RP_SLAB][RP_ROWS]>(park_raw);  // [4][32][32] x 16 B = 64 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t mt = blockIdx.x / (BM / RP_ROWS);
  const int r0 = (int)(blockIdx.x % (BM / RP_ROWS)) * RP_ROWS;
  const int64_t m0 = mt * BM + r0;
  const bool vec = (K % 4 == 0) && (((uintptr_t)x & 15) == 0) && (!z || ((uintptr_t)z & 15) == 0);
  {
    // pass 1: the warp's four rows side by side (eight independent 16-byte loads per k step and lane:
    // the pass is latency-bound, what matters is the number of loads in flight)
    constexpr int RW = RP_ROWS / (RP_THREADS / 32);  // rows per warp
    float ma[RW], ms[RW];
#pragma unroll
    for (int i = 0; i < RW; ++i) ma[i] = ms[i] = 0.f;
    if (vec) {
      for (int k = lane * 4; k < K; k += 128) {
        float4 a[RW], b[RW];
#pragma unroll
        for (int i = 0; i < RW; ++i) {
          const int64_t m = m0 + warp * RW + i;
          a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          b[i] = make_float4(1.f, 1.f, 1.f, 1.f);
          if (m < M) {
            a[i] = __ldg(reinterpret_cast<const float4*>(x + m * K + k));
            if (z) b[i] = __ldg(reinterpret_cast<const float4*>(z + m * K + k));
          }
        }
#pragma unroll
        for (int i = 0; i < RW; ++i) {
          ma[i] = fmaxf(ma[i], fmaxf(fmaxf(fabsf(a[i].x * b[i].x), fabsf(a[i].y * b[i].y)), fmaxf(fabsf(a[i].z * b[i].z), fabsf(a[i].w * b[i].w))));
          ms[i] = fmaxf(ms[i], fmaxf(fmaxf(a[i].x * a[i].x, a[i].y * a[i].y), fmaxf(a[i].z * a[i].z, a[i].w * a[i].w)));
        }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
#pragma unroll
        for (int i = 0; i < RW; ++i) {
          const int64_t m = m0 + warp * RW + i;
          if (m < M) {
            const float a = __ldg(x + m * K + k), b = z ? __ldg(z + m * K + k) : 1.f;
            ma[i] = fmaxf(ma[i], fabsf(a * b));
            ms[i] = fmaxf(ms[i], a * a);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RW; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ma[i] = fmaxf(ma[i], __shfl_xor_sync(0xffffffffu, ma[i], o));
        ms[i] = fmaxf(ms[i], __shfl_xor_sync(0xffffffffu, ms[i], o));
      }
      if (lane == 0) {
        const int rr = warp * RW + i;
        const int ea = scale_exp(ma[i]), es = scale_exp(ms[i]);
        fsc[rr][0] = ldexpf(1.f, ea);
        fsc[rr][1] = ldexpf(1.f, es);
        rsc[m0 + rr] = make_float2(ldexpf(1.f, -ea), ldexpf(1.f, -es));
      }
    }
  }
  __syncthreads();
  const int nchunks = kblocks * CH;
  for (int c0 = 0; c0 < nchunks; c0 += RP_SLAB) {
    const int c = c0 + lane, k0 = c * 8;  // this lane's chunk of the slab
    constexpr int RW = RP_ROWS / (RP_THREADS / 32);
    float4 xa[RW][2], za[RW][2];
    const bool fastk = vec && k0 + 8 <= K;
#pragma unroll
    for (int i = 0; i < RW; ++i) {
      const int64_t m = m0 + warp * RW + i;
      xa[i][0] = xa[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      za[i][0] = za[i][1] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (m < M && k0 < K) {
        const float* xr = x + m * K + k0;
        const float* zr = z ? z + m * K + k0 : nullptr;
        if (fastk) {
          xa[i][0] = __ldg(reinterpret_cast<const float4*>(xr));
          xa[i][1] = __ldg(reinterpret_cast<const float4*>(xr + 4));
          if (zr) {
            za[i][0] = __ldg(reinterpret_cast<const float4*>(zr));
            za[i][1] = __ldg(reinterpret_cast<const float4*>(zr + 4));
          }
        } else {
          float xv[8], zv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            xv[e] = 0.f; zv[e] = 1.f;
            if (k0 + e < K) {
              xv[e] = __ldg(xr + e);
              if (zr) zv[e] = __ldg(zr + e);
            }
          }
          xa[i][0] = make_float4(xv[0], xv[1], xv[2], xv[3]); xa[i][1] = make_float4(xv[4], xv[5], xv[6], xv[7]);
          za[i][0] = make_float4(zv[0], zv[1], zv[2], zv[3]); za[i][1] = make_float4(zv[4], zv[5], zv[6], zv[7]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RW; ++i) {
      const int rr = warp * RW + i;
      const float xv[8] = {xa[i][0].x, xa[i][0].y, xa[i][0].z, xa[i][0].w, xa[i][1].x, xa[i][1].y, xa[i][1].z, xa[i][1].w};
      const float zv[8] = {za[i][0].x, za[i][0].y, za[i][0].z, za[i][0].w, za[i][1].x, za[i][1].y, za[i][1].z, za[i][1].w};
      const float fa = fsc[rr][0], fs = fsc[rr][1];
      uint32_t ah[4], al[4], sh[4], sl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        split_f16x2(xv[2 * e] * zv[2 * e] * fa, xv[2 * e + 1] * zv[2 * e + 1] * fa, ah[e], al[e]);
        split_f16x2(xv[2 * e] * xv[2 * e] * fs, xv[2 * e + 1] * xv[2 * e + 1] * fs, sh[e], sl[e]);
      }
      const int slot = (rr ^ lane) & (RP_ROWS - 1);
      park[0][lane][slot] = make_uint4(ah[0], ah[1], ah[2], ah[3]);
      park[1][lane][slot] = make_uint4(al[0], al[1], al[2], al[3]);
      park[2][lane][slot] = make_uint4(sh[0], sh[1], sh[2], sh[3]);
      park[3][lane][slot] = make_uint4(sl[0], sl[1], sl[2], sl[3]);
    }
    __syncthreads();
    // copy-out: 32 / RP_ROWS (variant, chunk) lines per warp pass; a lane = one destination row of its line
    constexpr int LPW = 32 / RP_ROWS;
    const int sub = lane / RP_ROWS, row = lane % RP_ROWS;
    for (int lp = warp; lp < 4 * RP_SLAB / LPW; lp += RP_THREADS / 32) {
      const int ln = lp * LPW + sub;
      const int v = ln / RP_SLAB, cl = ln - v * RP_SLAB, cg = c0 + cl;
      if (cg < nchunks) {
        const int kb = cg / CH, cc = cg - kb * CH;
        uint8_t* dst = Ap + (mt * kblocks + kb) * (int64_t)A_STAGE + (int64_t)v * A_VAR + ((int64_t)cc * BM + r0 + row) * 16;
        *reinterpret_cast<uint4*>(dst) = park[v][cl][(row ^ cl) & (RP_ROWS - 1)];
      }
    }
    __syncthreads();
  }
}

// ---- weights: column maxima -> power-of-two scales (esc: exponent, wsc: inverse scale), clipped bias variance
__global__ void col_scale(int K, int N, int Npad, const float* __restrict__ Wm, const float* __restrict__ Ls, float max_std,
                          int* __restrict__ esc, float* __restrict__ wsc, const float* __restrict__ lvb,
                          float* __restrict__ vb) {
  // block = 32 columns x 8 k-groups (lanes walk columns: coalesced)
  __shared__ float red[8][32];
  const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + c, ph = blockIdx.y;
  float m = 0.f;
  if (n < N)
    for (int k = g; k < K; k += 8) {
      float w;
      if (ph == 0) w = fabsf(Wm[(int64_t)k * N + n]);
      else {
        const float sd = fminf(expf(Ls[(int64_t)k * N + n]), max_std);
        w = sd * sd;
      }
      m = fmaxf(m, w);
    }
  red[g][c] = m;
  __syncthreads();
  if (g == 0 && n < Npad) {
#pragma unroll
    for (int q = 1; q < 8; ++q) m = fmaxf(m, red[q][c]);
    const int e = n < N ? scale_exp(m) : 0;
    esc[ph * Npad + n] = e;
    wsc[ph * Npad + n] = ldexpf(1.f, -e);
    if (ph == 0 && n < N) vb[n] = fminf(expf(lvb[n]), max_std * max_std);
  }
}

// one thread = one 16-byte chunk (8 consecutive k of one output column) of all four variants
__global__ void weight_pack(int K, int N, int Npad, int kblocks, int ntiles, const float* __restrict__ Wm,
                            const float* __restrict__ Ls, float max_std, const int* __restrict__ esc,
                            uint8_t* __restrict__ Bp) {
  const int64_t total = (int64_t)ntiles * kblocks * (CH * BN);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx % BN);  // consecutive threads read consecutive columns of a weight row
    int64_t t = idx / BN;
    const int c = (int)(t % CH);
    t /= CH;
    const int kb = (int)(t % kblocks);
    const int nt = (int)(t / kblocks);
    const int gn = nt * BN + n, k0 = kb * KS + c * 8;
    float w[8], v[8];
    const int em = esc[gn], ev = esc[Npad + gn];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      w[e] = v[e] = 0.f;
      if (gn < N && k0 + e < K) {
        w[e] = ldexpf(__ldg(Wm + (int64_t)(k0 + e) * N + gn), em);
        const float s = fminf(expf(__ldg(Ls + (int64_t)(k0 + e) * N + gn)), max_std);
        v[e] = ldexpf(s * s, ev);
      }
    }
    uint32_t wh[4], wl[4], vh[4], vl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      split_f16x2(w[2 * e], w[2 * e + 1], wh[e], wl[e]);
      split_f16x2(v[2 * e], v[2 * e + 1], vh[e], vl[e]);
    }
    uint8_t* dst = Bp + ((int64_t)nt * kblocks + kb) * (int64_t)B_STAGE + ((int64_t)c * BN + n) * 16;
    *reinterpret_cast<uint4*>(dst) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
    *reinterpret_cast<uint4*>(dst + B_VAR) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    *reinterpret_cast<uint4*>(dst + 2 * B_VAR) = make_uint4(vh[0], vh[1], vh[2], vh[3]);
    *reinterpret_cast<uint4*>(dst + 3 * B_VAR) = make_uint4(vl[0], vl[1], vl[2], vl[3]);
  }
}

// INJ: the output noise is read from memory (tests); compiled apart so that the epilogue's Philox path is branch-free
// SEG: one K segment of a long contraction (a.seg = 1 .. 3, partial sums through a.raw); the ordinary whole-K launch
// is compiled without any of it
template <bool INJ, bool SEG>
__global__ void __launch_bounds__(THREADS, 1) paired_f16x3(const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;                    // [STAGES][4][CH][BM][16 B]
  uint8_t* Bs = As + STAGES * A_STAGE;   // [STAGES][4][CH][BN][16 B]
  uint64_t* full = (uint64_t*)(Bs + STAGES * B_STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;   // [2]
  uint64_t* tempty = tfull + 2;       // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int64_t ntot = a.mtiles * a.ntiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // tile order: the n-tiles of one m-tile are adjacent, so the CTAs working at the same time share
  // A tiles (and all of B, 3.2 MB) through L2
  if (warp < EPI_WARPS) {
    // ================================================================== epilogue warps
    const int r = warp * 32 + lane;  // tile row == TMEM lane
    const PhiloxKeys EK = philox_keys(a.eps);  // round keys once per warp (warp-uniform: uniform registers)
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntot; t += gridDim.x, ++it) {
      const int64_t mt = t / a.ntiles;
      const int nt = (int)(t - mt * a.ntiles);
      const int64_t m = mt * BM + r;
      const int n0 = nt * BN;
      const int buf = it & 1;
      const bool row_ok = m < a.M;
      const float2 rs = __ldg(a.rsc + mt * BM + r);  // before the wait: the load overlaps the MMAs
      const float* wm = a.wsc + n0;
      const float* wv = a.wsc + (int64_t)a.ntiles * BN + n0;
      mbar_wait(&tfull[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 2 * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        if (n0 + c0 >= a.N) break;  // warp-uniform
        float mv[16], vv[16];
        tmem_ld16(trow + c0, mv);
        tmem_ld16(trow + BN + c0, vv);
        if (c0 + 16 >= BN || n0 + c0 + 16 >= a.N) {  // last chunk: the accumulators are in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[buf]);
        }
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int nb = n0 + c0 + 4 * q;
            if (nb >= a.N) break;
            float e[4] = {0.f, 0.f, 0.f, 0.f};
            if (SEG && (a.seg == 1 || a.seg == 2)) {
            } else if (INJ) {
#pragma unroll
              for (int j = 0; j < 4; ++j) e[j] = nb + j < a.N ? a.eps.injected[m * a.N + nb + j] : 0.f;
            } else {
              philox_normal4_keys(a.eps, EK, a.eps.row_offset + (uint64_t)m, 0u, (uint32_t)(nb >> 2), e);
            }
            float o4[4] = {0.f, 0.f, 0.f, 0.f}, s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = nb + j, cl = c0 + 4 * q + j;
              if (n < a.N) {
                float mean, var;
                if (!SEG) {
                  mean = fmaf(mv[4 * q + j], rs.x * __ldg(wm + cl), __ldg(a.bm + n));
                  var = fmaf(vv[4 * q + j], rs.y * __ldg(wv + cl), __ldg(a.vb + n));
                } else {  // round-to-nearest fp32 sums of the segments' partial results
                  float pm = mv[4 * q + j] * (rs.x * __ldg(wm + cl)), pv = vv[4 * q + j] * (rs.y * __ldg(wv + cl));
                  float* rm = a.raw + m * a.N + n;
                  float* rv = rm + a.M * a.N;
                  if (a.seg >= 2) { pm += *rm; pv += *rv; }
                  if (a.seg <= 2) { *rm = pm; *rv = pv; }
                  mean = pm + __ldg(a.bm + n);
                  var = pv + __ldg(a.vb + n);
                }
                const float sd = sqrt_fast(var);
                s4[j] = sd;
                const float out = fmaf(sd, e[j], mean);
                o4[j] = a.act == 1 ? fmaxf(out, 0.f) : out;
              }
            }
            float* dstp = a.y + m * a.N + nb;
            if (SEG && (a.seg == 1 || a.seg == 2)) continue;  // partial sums only: the last segment writes the layer's output
            if (nb + 3 < a.N && !(a.N & 3)) {
              *reinterpret_cast<float4*>(dstp) = make_float4(o4[0], o4[1], o4[2], o4[3]);
              if (a.sd_out) *reinterpret_cast<float4*>(a.sd_out + m * a.N + nb) = make_float4(s4[0], s4[1], s4[2], s4[3]);
            } else {
              for (int j = 0; j < 4; ++j)
                if (nb + j < a.N) {
                  dstp[j] = o4[j];
                  if (a.sd_out) a.sd_out[m * a.N + nb + j] = s4[j];
                }
            }
          }
        }
      }
    }
  } else if (warp == EPI_WARPS) {
    // ======================================================================= MMA issuer
    // instruction descriptor: D = F32 (bit 4), A = B = F16 (format 0), K-major both, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint64_t adesc0 = make_desc(smem_u32(As), BM * 16, 128), bdesc0 = make_desc(smem_u32(Bs), BN * 16, 128);
    int it = 0, stage = 0;
    uint32_t sph = 0;
    for (int64_t t = blockIdx.x; t < ntot; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_mean = tmem_base + (uint32_t)(buf * 2 * BN), d_var = d_mean + BN;
      for (int kb = a.kb0; kb < a.kb1; ++kb) {
        mbar_wait(&full[stage], sph);
        tc_fence_after();
        const uint64_t ad = adesc0 + (uint64_t)((stage * A_STAGE) >> 4);
        const uint64_t bd = bdesc0 + (uint64_t)((stage * B_STAGE) >> 4);
#pragma unroll
        for (int j = 0; j < KS / 16; ++j) {  // one K = 16 instruction = two 16-byte chunks
          const uint64_t ao = ad + (uint64_t)((2 * j * BM * 16) >> 4), bo = bd + (uint64_t)((2 * j * BN * 16) >> 4);
          const uint32_t acc = (kb == a.kb0 && j == 0) ? 0u : 1u;
          constexpr uint64_t AV = A_VAR >> 4, BV = B_VAR >> 4;
          umma_f16_elect(d_mean, ao, bo, idesc, acc);            // xz_hi * W_hi
          umma_f16_elect(d_mean, ao + AV, bo, idesc, 1u);        // xz_lo * W_hi
          umma_f16_elect(d_mean, ao, bo + BV, idesc, 1u);        // xz_hi * W_lo
          umma_f16_elect(d_var, ao + 2 * AV, bo + 2 * BV, idesc, acc);   // s_hi * V_hi
          umma_f16_elect(d_var, ao + 3 * AV, bo + 2 * BV, idesc, 1u);    // s_lo * V_hi
          umma_f16_elect(d_var, ao + 2 * AV, bo + 3 * BV, idesc, 1u);    // s_hi * V_lo
        }
        umma_commit_elect(&empty[stage]);                         // smem stage free once these MMAs retire
        if (kb == a.kb1 - 1) umma_commit_elect(&tfull[buf]);  // accumulators ready
        __syncwarp();
        if (++stage == STAGES) { stage = 0; sph ^= 1; }
      }
    }
  } else {
    // ======================================================================== bulk-copy producer
    int stage = 0;
    uint32_t sph = 0;
    for (int64_t t = blockIdx.x; t < ntot; t += gridDim.x) {
      const int64_t mt = t / a.ntiles;
      const int nt = (int)(t - mt * a.ntiles);
      const uint8_t* asrc = a.Ap + mt * a.kblocks * (int64_t)A_STAGE;
      const uint8_t* bsrc = a.Bp + (int64_t)nt * a.kblocks * (int64_t)B_STAGE;
      for (int kb = a.kb0; kb < a.kb1; ++kb) {
        mbar_wait(&empty[stage], sph ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full[stage], A_STAGE + B_STAGE);
          bulk_g2s(As + stage * A_STAGE, asrc + (int64_t)kb * A_STAGE, A_STAGE, &full[stage]);
          bulk_g2s(Bs + stage * B_STAGE, bsrc + (int64_t)kb * B_STAGE, B_STAGE, &full[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; sph ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

}  // namespace d3

// K beyond MNF_TC_PARITY_KMAX: segments of SEG_KB k-blocks (1024 terms per TMEM accumulation chain), one GEMM launch each
constexpr int SEG_KB = 1024 / d3::KS;
bool mnf_dense_f16x3_ok(const mnf_ctx* ctx, int64_t M, int K, int N) {
  return ctx->math_mode == MNF_MATH_TF32X3 && M >= 1024 && N > 16 && K >= 64 && K <= 16384 && M * (int64_t)N < (1ll << 31);
}

// Entry used by layers_fwd.cu for many-row dense layers in the fp32-parity tensor-core mode.
int mnf_dense_f16x3_forward(mnf_ctx* ctx, int64_t M, int K, int N, const float* x, const float* z, const float* mean_W,
                            const float* log_std_W, const float* mean_b, const float* log_var_b, float max_std,
                            const NoiseDev& eps, float* y, float* sd_out, int act) {
  using namespace d3;
  Args a = {};
  a.M = M; a.K = K; a.N = N;
  a.mtiles = cdiv(M, BM);
  a.kblocks = (K + KS - 1) / KS;
  a.ntiles = (N + BN - 1) / BN;
  a.bm = mean_b; a.y = y; a.sd_out = sd_out; a.eps = eps; a.act = act;
  const int Npad = a.ntiles * BN;
  const size_t a_bytes = (size_t)a.mtiles * a.kblocks * A_STAGE;
  const size_t r_bytes = sizeof(float2) * (size_t)a.mtiles * BM;
  const size_t bp_bytes = (size_t)a.ntiles * a.kblocks * B_STAGE;
  const size_t b_bytes = bp_bytes + sizeof(float) * (size_t)(2 * Npad + Npad) + sizeof(int) * (size_t)(2 * Npad);
  const int nseg = K > MNF_TC_PARITY_KMAX ? (a.kblocks + SEG_KB - 1) / SEG_KB : 1;
  const size_t raw_bytes = nseg > 1 ? (sizeof(float) * 2 * (size_t)M * N + 255) & ~(size_t)255 : 0;
  uint8_t *Ap, *Bp;
  bool fresh = true;
  if (ctx->frozen) {
    // the packed weights persist (keyed by the parameter buffers); only the rows are packed per call
    Bp = (uint8_t*)mnf_packed(ctx, mnf_mix(0xF16D3ull, ((uint64_t)K << 32) | (uint32_t)N), (uint64_t)(uintptr_t)mean_W,
                              (uint64_t)(uintptr_t)log_std_W,
                              mnf_mix((uint64_t)(uintptr_t)log_var_b, (uint64_t)__float_as_uint_host(max_std)), b_bytes, &fresh);
    if (!Bp) return MNF_ERR_CUDA;
    Ap = (uint8_t*)mnf_workspace(ctx, a_bytes + r_bytes + raw_bytes);
    if (!Ap) return MNF_ERR_CUDA;
    a.raw = (float*)(Ap + a_bytes + r_bytes);
  } else {
    Ap = (uint8_t*)mnf_workspace(ctx, a_bytes + r_bytes + raw_bytes + b_bytes);
    if (!Ap) return MNF_ERR_CUDA;
    a.raw = (float*)(Ap + a_bytes + r_bytes);
    Bp = Ap + a_bytes + r_bytes + raw_bytes;
  }
  float2* rsc = (float2*)(Ap + a_bytes);
  float* wsc = (float*)(Bp + bp_bytes);
  float* vb = wsc + 2 * Npad;
  int* esc = (int*)(vb + Npad);
  if (fresh) {
    col_scale<<<dim3((unsigned)(Npad / 32), 2), 256, 0, ctx->stream>>>(K, N, Npad, mean_W, log_std_W, max_std, esc, wsc, log_var_b, vb);
    MNF_LAUNCHED(ctx);
    const int64_t tot = (int64_t)a.ntiles * a.kblocks * (CH * BN);
    const int grid = (int)(cdiv(tot, 256) < (int64_t)ctx->sm_count * 16 ? cdiv(tot, 256) : (int64_t)ctx->sm_count * 16);
    weight_pack<<<grid, 256, 0, ctx->stream>>>(K, N, Npad, a.kblocks, a.ntiles, mean_W, log_std_W, max_std, esc, Bp);
    MNF_LAUNCHED(ctx);
  }
  const bool probed = mnf_probe_begin(ctx, MNF_PROBE_DENSE_GEMM);  // the layer's two per-call launches: pack + GEMM
  static bool pack_attr = false;
  if (!pack_attr) {
    MNF_CUDA(cudaFuncSetAttribute(row_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, RP_PARK_BYTES));
    pack_attr = true;
  }
  row_pack<<<(unsigned)(a.mtiles * (BM / RP_ROWS)), RP_THREADS, RP_PARK_BYTES, ctx->stream>>>(M, K, a.kblocks, x, z, Ap, rsc);
  MNF_LAUNCHED(ctx);
  a.Ap = Ap; a.Bp = Bp; a.rsc = rsc; a.wsc = wsc; a.vb = vb;
  const size_t smem = (size_t)STAGES * (A_STAGE + B_STAGE) + 256;
  static bool attr_set = false;  // one device per process (mnf_ctx_create)
  if (!attr_set) {
    MNF_CUDA(cudaFuncSetAttribute(paired_f16x3<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MNF_CUDA(cudaFuncSetAttribute(paired_f16x3<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MNF_CUDA(cudaFuncSetAttribute(paired_f16x3<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MNF_CUDA(cudaFuncSetAttribute(paired_f16x3<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t tiles = a.mtiles * a.ntiles;
  const int g = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
  mnf_note_path(nseg > 1 ? "paired_f16x3<128>/segments" : "paired_f16x3<128>");
  for (int sg = 0; sg < nseg; ++sg) {
    a.kb0 = sg * (nseg > 1 ? SEG_KB : a.kblocks);
    a.kb1 = nseg > 1 ? (a.kb0 + SEG_KB < a.kblocks ? a.kb0 + SEG_KB : a.kblocks) : a.kblocks;
    a.seg = nseg == 1 ? 0 : sg == 0 ? 1 : sg == nseg - 1 ? 3 : 2;
    if (nseg > 1) {
      if (eps.injected) paired_f16x3<true, true><<<g, THREADS, smem, ctx->stream>>>(a);
      else paired_f16x3<false, true><<<g, THREADS, smem, ctx->stream>>>(a);
    } else {
      if (eps.injected) paired_f16x3<true, false><<<g, THREADS, smem, ctx->stream>>>(a);
      else paired_f16x3<false, false><<<g, THREADS, smem, ctx->stream>>>(a);
    }
    MNF_LAUNCHED(ctx);
  }
  mnf_probe_end(ctx, probed);
  return MNF_OK;
}
