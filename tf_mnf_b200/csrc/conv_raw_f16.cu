// MNFConv2D.call (mnf_conv.py:182-197) on tcgen05 kind::f16 with split operands -- the raw-tile
// kernel of conv_raw_f16.cu (im2col expressed by per-MMA UMMA descriptors over the staged input
// tile) with HALF the bytes per K element.
//
// conv_raw_f16.cu is bound by the L2 -> SM stream of its weights (every 128-row tile re-reads
// all K x 64 x {hi, lo} x {mean, var} TF32 weights: 2.6 GB per MNF-LeNet conv2 launch) and by
// the shared-memory operand fetch of its K=8 MMAs.  fp16 carries the same 11 significant bits as
// TF32 in 2 bytes, so the same hi/lo split (x = hi + lo, x*w ~ hi*hi + lo*hi + hi*lo, fp32
// accumulate) needs half the weight bytes and half the MMAs (K = 16 per instruction).  What
// fp16 lacks is range, so every operand gets an exact power-of-two scale:
//   weights: per output column n, 2^e so that max_k |w[k,n]| * 2^e lies in [2^13, 2^14)
//            (mean and variance weights separately), found by a column-max pre-pass;
//   inputs:  per image of the tile, 2^e so that max |x| * 2^e lies in [2^6, 2^7), hence
//            (x * 2^e)^2 < 2^14; found by the producer warps while the previous tile computes.
// The epilogue multiplies the fp32 accumulators by the inverse scales (exact).  Elements far
// below their column / image maximum lose RELATIVE precision in the lo part (fp16 subnormals,
// quantum 2^-24 of the scaled value) but their absolute error stays below 2^-38 of the maximum,
// far inside the 1e-5 parity bar, which is relative to the output scale.
//
//   planes[v][cc][hi][img][wi] of 16-byte chunks (8 channels, zero padded), v in {x hi, x lo,
//   x^2 hi, x^2 lo};  step table, phases, warp roles and the N=128 [W_hi | W_lo] instruction as
//   in conv_raw_f16.cu.
#include <cuda_fp16.h>
#include <cstdlib>

#include <utility>
#include <vector>

#include "umma.cuh"

namespace cf {
using namespace um;

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u));  // no "memory" clobber: volatile orders it against
}                                                              // the (volatile) barrier waits and commits
// (a0, a1) -> packed fp16 hi pair and lo pair, a = hi + lo (+ O(2^-22 |a|)), both rounded to nearest
__device__ __forceinline__ void split_f16x2(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  // packed conversions (F2FP.F16.F32.PACK_AB: one instruction per PAIR, not on the quarter-rate XU pipe that the
  // scalar F2F.F16.F32 uses -- the plane producers convert 4 values per element)
  const __half2 h = __floats2half2_rn(a0, a1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

constexpr int BM = 128, BN = 64;
constexpr int G = 6;                 // K=16 steps per weight stage
constexpr int STEP_B = 2 * 2 * BN * 16;  // bytes of one K=16 step's weights for one phase: [2 kc][hi 64 n | lo 64 n][8 halves] = 4 KB
constexpr int B_STAGE = G * STEP_B;  // 24 KB
constexpr int SB = 3;                // weight ring depth
constexpr int THREADS = 20 * 32;
constexpr int MAXSTEPS = 64;
constexpr int TT = 2;              // tiles per pass over the weights (halves the L2 -> SM weight stream)
constexpr int MAXC = 3;            // plane chunks a producer thread keeps in registers
constexpr int PROD_WARPS = 10;     // warps 5-7 and 13-19 fill the planes (the fills pace the MMAs); warps 8-11: second epilogue group
constexpr int TMEM_COLS = 512;  // [2 buffers][TT tiles][2 phases][64], double buffered over passes

struct Args {
  int64_t B;         // images
  int H, W, C, kh, kw, Ho, N, I;  // Wo == 8; I images per tile
  int Hi, Wi;        // plane extent per image: Ho-1+kh, 8-1+kw
  int half;          // C % 8 == 4: the last chunk of a pixel carries 4 channels of this pixel and 4 of the next
  int cpt;           // 16-byte channel chunks per pixel (ceil(C / 8))
  int steps;         // K=16 steps
  int pool;
  const float* x;
  const float* z;    // [B,N] or nullptr
  const float* bm;
  const float* vb;
  const int* steptab;  // [steps][2]: a_off, a_lbo (bytes)
  const float* Bp;     // [2 phases][steps][STEP_B bytes] fp16 pairs
  const float* wsc;    // [2][BN] inverse column scales of the mean / variance weights
  float* y;
  float* mean_out;
  float* sd_out;
  NoiseDev eps;
  uint64_t dtab[MAXSTEPS];  // A descriptor of every step without the tile base: constant bank -> uniform registers
};

__global__ void __launch_bounds__(THREADS, 1) conv_raw_f16(const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int pix = a.Hi * a.I * a.Wi;          // 16-byte chunks per channel-chunk plane
  const int CCS = pix * 16;                   // bytes between channel chunks
  const int PLANE = a.cpt * CCS;              // bytes of one variant
  uint8_t* planes = smem;                     // [TT][4][cpt][Hi][I][Wi][16 B]
  uint8_t* Bs = planes + TT * 4 * PLANE;      // [SB][B_STAGE]
  float* ys = (float*)(Bs + SB * B_STAGE);    // [BM][N] dense staging
  int* stab = (int*)(ys + BM * BN);              // [steps][2]
  float* bias_s = (float*)(stab + 2 * a.steps);  // [2][BN]
  float* zs = bias_s + 2 * BN;                   // [I][BN] z rows of the tile being written out
  float* wsc_s = zs + 16 * BN;                   // [2][BN] inverse weight-column scales
  float* xsc = wsc_s + 2 * BN;                   // [4 passes][TT][16] inverse input scale per image
  float* xsf = xsc + 128;                        // [TT][16] forward input scale per image (producers)
  unsigned* xmax = (unsigned*)(xsf + 32);        // [TT][16] max |x| per image (float bits)
  uint64_t* bars = (uint64_t*)(((uintptr_t)(xmax + 32) + 7) & ~(uintptr_t)7);
  uint64_t* pfull = bars;                 // [TT][2] planes of (tile, phase) ready
  uint64_t* pempty = pfull + 2 * TT;      // [TT][2]
  uint64_t* bfull = pempty + 2 * TT;      // [SB]
  uint64_t* bempty = bfull + SB;          // [SB]
  uint64_t* tfull = bempty + SB;          // [2][TT] both accumulators of a tile complete
  uint64_t* tempty = tfull + 2 * TT;      // [2][TT]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2 * TT);

  // (Measured: moving the MMA issuer to the highest hardware warp id, which the sub-partition's arbiter prefers,
  // made this kernel 13 % SLOWER -- the producers, not the issuer, need the issue slots.)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int HoWo = a.Ho * 8;
  const int64_t ntiles = (a.B + a.I - 1) / a.I;
  const int64_t npass = (ntiles + TT - 1) / TT;   // a pass = TT tiles sharing one stream of the weights
  const int nstages = (a.steps + G - 1) / G;

  if (threadIdx.x == 0) {
    for (int p = 0; p < 2 * TT; ++p) { mbar_init(&pfull[p], PROD_WARPS); mbar_init(&pempty[p], 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int b = 0; b < 2 * TT; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8); }
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
    wsc_s[n] = a.wsc[n];
    wsc_s[BN + n] = a.wsc[BN + n];
  }
  if (threadIdx.x < 32) xmax[threadIdx.x] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler

  if (warp < 4 || (warp >= 8 && warp < 12)) {
    // ================================================================== epilogue warps
    const int quad = warp & 3, chalf = warp >= 8 ? 1 : 0;
    const int et = chalf * 128 + quad * 32 + lane;   // 0..255 among the epilogue threads
    const int r = quad * 32 + lane;             // TMEM lane = tile row (ho, img, wo)
    const int wo = r & 7, g = r >> 3, img = g % a.I, ho = g / a.I;
    const int rl = img * HoWo + ho * 8 + wo;    // row inside the tile's block of y
    // Fast path (inference with the fused pool, I <= 2 -- MNF-LeNet conv2): the four pixels of a pool
    // window sit in one warp (wo pair: lane ^ 1, ho pair: lane ^ 8*I), so the max is two shuffles and
    // the pooled value is stored straight from registers: no staging tile, no block-wide barriers.
    const bool fastpool = a.pool && a.I <= 2 && !a.sd_out && !a.mean_out && !(a.N & 1);
    if (fastpool) {
      const int hbit = 8 * a.I;
      const bool writer = (lane & 1) == 0 && (lane & hbit) == 0;
      const int Hp = a.Ho >> 1;
      int it = 0;
      for (int64_t ps = blockIdx.x; ps < npass; ps += gridDim.x, ++it)
        for (int tl = 0; tl < TT; ++tl) {
          const int64_t braw = (ps * TT + tl) * a.I + img;
          const bool st = braw < a.B && writer;
          const int64_t brow = braw < a.B ? braw : a.B - 1;  // shuffles need every lane: compute on a valid row
          const int buf = it & 1;
          const uint32_t outer = (uint32_t)(ho * 8 + wo);
          const int64_t m = brow * HoWo + outer;
          const uint64_t grow = a.eps.row_offset + (uint64_t)brow;
          const float* zrow = a.z ? a.z + brow * a.N : nullptr;
          float* dst = a.y + ((brow * Hp + (ho >> 1)) * 4 + (wo >> 1)) * a.N;
          mbar_wait(&tfull[buf * TT + tl], (it >> 1) & 1);
          tc_fence_after();
          const float sx = xsc[((it & 3) * TT + tl) * 16 + img], sx2 = sx * sx;
          const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((buf * TT + tl) * 2 * BN);
#pragma unroll 1
          for (int c0 = chalf * 32; c0 < chalf * 32 + 32; c0 += 16) {
            if (c0 >= a.N) break;
            float mv[16], vv[16];
            tmem_ld16(trow + c0, mv);
            tmem_ld16(trow + BN + c0, vv);
            if (c0 + 16 >= a.N || c0 + 16 >= chalf * 32 + 32) {  // last chunk of this warp: accumulators are in registers
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[buf * TT + tl]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int nb = c0 + 4 * q;
              if (nb >= a.N) break;
              float e[4];
              if (a.eps.injected) {
#pragma unroll
                for (int j = 0; j < 4; ++j) e[j] = nb + j < a.N ? a.eps.injected[m * a.N + nb + j] : 0.f;
              } else {
                philox_normal4(a.eps, grow, outer, (uint32_t)(nb >> 2), e);
              }
              float o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int n = nb + j < a.N ? nb + j : a.N - 1;
                float mean = mv[4 * q + j] * (sx * wsc_s[n]) + bias_s[n];
                const float sd = sqrt_fast(vv[4 * q + j] * (sx2 * wsc_s[BN + n]) + bias_s[BN + n]);
                if (zrow) mean *= __ldg(zrow + n);
                float v = fmaf(sd, e[j], mean);
                v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));     // wo pair
                v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, hbit));  // ho pair
                o[j] = fmaxf(v, 0.f);
              }
              if (st) {
                *reinterpret_cast<float2*>(dst + nb) = make_float2(o[0], o[1]);
                if (nb + 2 < a.N) *reinterpret_cast<float2*>(dst + nb + 2) = make_float2(o[2], o[3]);
              }
            }
          }
        }
    } else {
    int it = 0;
    for (int64_t ps = blockIdx.x; ps < npass; ps += gridDim.x, ++it)
    for (int tl = 0; tl < TT; ++tl) {
      const int64_t t = ps * TT + tl;
      const int64_t b0 = t * a.I, brow = b0 + img;
      // this tile's z rows -> shared memory (the previous tile's readers passed the closing barrier)
      for (int e = et; e < a.I * BN; e += 256) {
        const int im = e / BN, n = e - im * BN;
        zs[e] = (a.z && b0 + im < a.B && n < a.N) ? __ldg(a.z + (b0 + im) * a.N + n) : 1.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int buf = it & 1;
      mbar_wait(&tfull[buf * TT + tl], (it >> 1) & 1);
      tc_fence_after();
      const bool row_ok = brow < a.B;
      const float sx = xsc[((it & 3) * TT + tl) * 16 + img], sx2 = sx * sx;  // written before the planes were filled
      const uint32_t outer = (uint32_t)(ho * 8 + wo);
      const int64_t m = brow * HoWo + outer;
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((buf * TT + tl) * 2 * BN);
#pragma unroll 1
      for (int c0 = chalf * 32; c0 < chalf * 32 + 32; c0 += 16) {
        if (c0 >= a.N) break;
        float mv[16], vv[16];
        tmem_ld16(trow + c0, mv);
        tmem_ld16(trow + BN + c0, vv);
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
                float mean = mv[4 * q + j] * (sx * wsc_s[n]) + bias_s[n];
                const float sd = sqrt_fast(vv[4 * q + j] * (sx2 * wsc_s[BN + n]) + bias_s[BN + n]);
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
      if (lane == 0) mbar_arrive(&tempty[buf * TT + tl]);
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
    }
  } else if (warp == 4) {
    // ======================================================================= MMA issuer
    // three N = 64 instructions per step and phase (x_hi*W_hi, x_lo*W_hi, x_hi*W_lo) into the same
    // 64 TMEM columns; see the comment at the issue site for why not the N = 128 [W_hi | W_lo] form
    // instruction descriptor: D = F32 (bit 4), A = B = F16 (format 0), K-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t sbo = (uint32_t)(a.Wi * 16);
    const uint64_t bdesc0 = make_desc(smem_u32(Bs), 2 * BN * 16, 128);
    const uint32_t plane0 = smem_u32(planes);
    int it = 0, st = 0;
    uint32_t bph = 0;
    for (int64_t ps = blockIdx.x; ps < npass; ps += gridDim.x, ++it) {
      const int buf = it & 1;
      for (int ph = 0; ph < 2; ++ph) {
        for (int sg = 0; sg < nstages; ++sg) {
          mbar_wait(&bfull[st], bph);
          for (int tl = 0; tl < TT; ++tl) {
            if (sg == 0) {
              if (ph == 0) mbar_wait(&tempty[buf * TT + tl], ((it >> 1) & 1) ^ 1);  // accumulators drained by the epilogue
              mbar_wait(&pfull[tl * 2 + ph], it & 1);
            }
            tc_fence_after();
            {
              // warp-converged issue (umma.cuh): per MMA only an add of the (tile, phase) base to
              // the step's pre-built descriptor, in uniform registers
              const uint32_t dacc = tmem_base + (uint32_t)((buf * TT + tl) * 2 * BN + ph * BN);
              const uint64_t abase = (uint64_t)((plane0 + (uint32_t)((tl * 4 + 2 * ph) * PLANE)) >> 4);
              const uint64_t bst = bdesc0 + (uint64_t)((st * B_STAGE) >> 4);
              const int s0 = sg * G, ns = min(G, a.steps - s0);
#pragma unroll
              for (int j = 0; j < G; ++j) {
                if (j < ns) {
                  const uint64_t ah = a.dtab[s0 + j] + abase, al = ah + (uint64_t)(PLANE >> 4);
                  const uint64_t bh = bst + (uint64_t)((j * STEP_B) >> 4);
                  // three N = 64 instructions into the same 64 columns (with one tile per weight pass
                  // the [W_hi | W_lo] N = 128 form is cheaper, but it needs twice the TMEM columns
                  // and two tiles per pass halve the L2 -> SM weight stream)
                  umma_f16_elect(dacc, ah, bh, idesc, (s0 + j) == 0 ? 0u : 1u);
                  umma_f16_elect(dacc, al, bh, idesc, 1u);
                  umma_f16_elect(dacc, ah, bh + (uint64_t)((BN * 16) >> 4), idesc, 1u);
                }
              }
              if (sg == nstages - 1) {
                umma_commit_elect(&pempty[tl * 2 + ph]);       // planes of this (tile, phase) may be refilled
                if (ph == 1) umma_commit_elect(&tfull[buf * TT + tl]);    // both accumulators of the tile complete
              }
              if (tl == TT - 1) umma_commit_elect(&bempty[st]);
            }
            __syncwarp();
          }
          if (++st == SB) { st = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 12) {
    // ==================================================================== weight loader
    if (lane == 0) {
      int st = 0;
      uint32_t bph = 0;
      for (int64_t ps = blockIdx.x; ps < npass; ps += gridDim.x) {
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
    if (warp < 8 || warp > 12) {
      const int pt = (warp < 8 ? warp - 5 : warp - 10) * 32 + lane;  // 0 .. 319
      constexpr int NP = PROD_WARPS * 32;
      const int nchunks = a.cpt * pix;       // 16-byte chunks per variant (<= MAXC * NP, host-checked)
      const int img_f = a.H * a.W * a.C;     // floats per input image
      // this thread's chunks: plane offset, image of the tile, float offset inside the image
      // (-1: outside the input -> zeros), whether channels 4..7 of the chunk exist
      int o[MAXC], gsrc[MAXC], imx[MAXC];
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int q = pt + c * NP;
        o[c] = -1; gsrc[c] = -1; imx[c] = 0;
        if (q < nchunks) {
          // q enumerates (hi, img, wi, cc), cc fastest: a warp reads consecutive floats
          const int cc = q % a.cpt;
          int p = q / a.cpt;
          const int wi = p % a.Wi;
          p /= a.Wi;
          const int im = p % a.I, hi = p / a.I;
          o[c] = cc * CCS + ((hi * a.I + im) * a.Wi + wi) * 16;
          imx[c] = im | (cc * 8 + 4 < a.C ? 256 : 0);
          if (hi < a.H && wi < a.W) gsrc[c] = (hi * a.W + wi) * a.C + cc * 8;
          if (a.half && cc == a.cpt - 1 && hi < a.H && wi + 1 < a.W) imx[c] |= 512;  // second half: next pixel
        }
      }
      int it = 0;
      for (int64_t ps = blockIdx.x; ps < npass; ps += gridDim.x, ++it) {
        // order: (tile 0, x) (tile 1, x) (tile 0, x^2) (tile 1, x^2) -- the x planes of BOTH tiles are
        // released half a pass before the x^2 planes, and the next pass starts with them
        for (int ph = 0; ph < 2; ++ph)
          for (int tl = 0; tl < TT; ++tl) {
            const int64_t b0 = (ps * TT + tl) * a.I;
            // ---- every chunk of the tile into registers (the second visit hits L2)
            float4 v[MAXC][2];
#pragma unroll
            for (int c = 0; c < MAXC; ++c) {
              v[c][0] = v[c][1] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (gsrc[c] >= 0) {
                int64_t b = b0 + (imx[c] & 255);
                if (b >= a.B) b = a.B - 1;
                const float* src = a.x + b * img_f + gsrc[c];
                v[c][0] = __ldg(reinterpret_cast<const float4*>(src));
                if (imx[c] & 256) v[c][1] = __ldg(reinterpret_cast<const float4*>(src + 4));
                else if (imx[c] & 512) v[c][1] = __ldg(reinterpret_cast<const float4*>(src + a.C));
              }
            }
            const float* sc = xsf + tl * 16;  // forward scale per image
            if (ph == 0) {
              // ---- max |x| per image -> power-of-two scales
              unsigned* xm = xmax + tl * 16;
              for (int im = 0; im < a.I; ++im) {
                float m = 0.f;
#pragma unroll
                for (int c = 0; c < MAXC; ++c)
                  if (o[c] >= 0 && (imx[c] & 255) == im) {
                    m = fmaxf(m, fmaxf(fmaxf(fabsf(v[c][0].x), fabsf(v[c][0].y)), fmaxf(fabsf(v[c][0].z), fabsf(v[c][0].w))));
                    m = fmaxf(m, fmaxf(fmaxf(fabsf(v[c][1].x), fabsf(v[c][1].y)), fmaxf(fabsf(v[c][1].z), fabsf(v[c][1].w))));
                  }
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sh));
                if (lane == 0) atomicMax(xm + im, __float_as_uint(m));
              }
              asm volatile("bar.sync 2, %0;" ::"n"(NP) : "memory");
              if (pt < a.I) {
                const float m = __uint_as_float(xm[pt]);
                int e = 0;
                if (m > 0.f && m < 3.0e38f) {
                  int ex;
                  frexpf(m, &ex);          // m = f * 2^ex, f in [0.5, 1)
                  e = 7 - ex;              // m * 2^e in [2^6, 2^7)
                  e = e > 100 ? 100 : (e < -100 ? -100 : e);
                }
                xsf[tl * 16 + pt] = ldexpf(1.f, e);
                xsc[((it & 3) * TT + tl) * 16 + pt] = ldexpf(1.f, -e);
                xm[pt] = 0u;  // ready for the next pass
              }
              asm volatile("bar.sync 2, %0;" ::"n"(NP) : "memory");
            }
            mbar_wait(&pempty[tl * 2 + ph], (it & 1) ^ 1);
            uint8_t* dst_hi = planes + (tl * 4 + 2 * ph) * PLANE;
            uint8_t* dst_lo = dst_hi + PLANE;
#pragma unroll
            for (int c = 0; c < MAXC; ++c) {
              if (o[c] < 0) continue;
              const float f = sc[imx[c] & 255];
              const float in[8] = {v[c][0].x, v[c][0].y, v[c][0].z, v[c][0].w, v[c][1].x, v[c][1].y, v[c][1].z, v[c][1].w};
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a0 = in[2 * e] * f, a1 = in[2 * e + 1] * f;
                if (ph == 1) { a0 *= a0; a1 *= a1; }
                split_f16x2(a0, a1, hw[e], lw[e]);
              }
              *reinterpret_cast<uint4*>(dst_hi + o[c]) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(dst_lo + o[c]) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pfull[tl * 2 + ph]);
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

// =====================================================================================================
// The same kernel on CTA PAIRS (tcgen05 cta_group::2): two SMs of one TPC run ONE M = 256 instruction per
// K = 16 step, each on its own 128-row tile and each holding HALF of the step's weight rows.
//
// Why: the single-CTA kernel above is bound by the shared-memory fetch of its SS-mode N = 64 MMAs -- A
// (128 rows x 32 B = 4 KB) is re-read by each of the three instructions of a product next to 2 KB of B:
// 52 cycles measured for a 32-cycle instruction.  It shares a weight stage between TWO tiles per CTA
// to halve the L2 -> SM weight stream.  A CTA pair gets both effects per SM: a weight stage is shared by
// the two tiles of the pair (one per SM) and every SM fetches only 32 of the 64 B rows per instruction
// (4 KB + 1 KB: 40 cycles), planes and TMEM per SM shrink to one tile, passes per SM double (better
// tail).  Roles per CTA as above with TT = 1; differences:
//   * only the leader CTA (cluster rank 0) issues MMAs; tcgen05.commit multicasts every "consumed" /
//     "accumulator ready" signal to the barrier at the same offset in BOTH CTAs (pempty, bempty, tfull);
//   * what the leader must wait for in the PEER CTA (planes filled, weight stage landed, accumulators
//     drained) is forwarded by the peer's otherwise idle warp 4: it waits on the peer's local barrier and
//     arrives remotely (mapa + mbarrier.arrive.release.cluster) on the leader's mirror barrier
//     (pfull_peer, bfull_peer, tempty_peer), in exactly the order the leader waits.
// Every wait is bounded (umma.cuh): a protocol error traps.
constexpr int STEP_B2 = 2 * 2 * (BN / 2) * 16;  // bytes of one K=16 step's weights for one phase in ONE CTA: [2 kc][hi 32 n | lo 32 n][8 halves]
constexpr int G2 = 11;                          // steps per weight stage (MNF-LeNet conv2: 33 steps = 3 full stages)
constexpr int B_STAGE2 = G2 * STEP_B2;          // 22 KB
constexpr int SB2 = 3;
constexpr int TMEM_COLS2 = 256;                 // [2 buffers][2 phases][64]
constexpr int THREADS2 = 20 * 32;               // 8 epilogue warps (0-3, 8-11), MMA 4, loader 12, 10 producers (5-7, 13-19)
constexpr int PROD_WARPS2 = 10;                 // (measured: 14 + 4 starves the epilogue, 12 + 8 on 22 warps loses registers)
constexpr int EPI_WARPS2 = 8;                   // TMEM lane quadrant x column half

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Waits stay at CTA scope (umma.cuh: mbar_wait).  A cluster-scope acquire on the polling side makes ptxas emit
// CCTL.IVALL -- an L1 invalidate -- behind EVERY try_wait (6 % of this kernel's instructions when it was
// tried, with the producers' and epilogue's global loads missing L1 ever after), and buys nothing here: what
// the peer CTA publishes is shared memory read by its own SM's tensor-core datapath, ordered by the
// arriving side's fence.proxy.async + release.cluster arrive and by tcgen05.commit.
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
// arrive on the barrier at the same offset in CTA `rank` of the cluster (cluster-scope release)
__device__ __forceinline__ void mbar_arrive_cl(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void umma2_f16_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void umma2_commit_mc_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// Diagnostic (MNF_CONV_PROF=1 prints it): cycles the leader's MMA warp of cluster 0 spent waiting per barrier
// kind [bfull, bfull_peer, tempty, -, pfull, -, -, total]
__device__ unsigned long long g_pair_prof[8];

// Barriers: every "ready" barrier the leader's MMA warp waits on lives in the LEADER CTA and counts the arrivals
// of BOTH CTAs -- the peer's producer / epilogue warps arrive on it remotely (mapa + cluster-scope release):
// pfull[set] (2 x 10 producer warps), tempty[buf] (2 x 8 epilogue warps).  The one exception is "weight stage
// landed": a bulk copy completes on a barrier of its destination CTA, so the peer's warp 4 (it has no MMAs to
// issue) waits on the peer's bfull[st] and forwards one arrival to the leader's bfull_peer[st].  Every
// "consumed" signal is a tcgen05.commit multicast to the same barrier offset in both CTAs (pempty, bempty, tfull).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS2, 1) conv_raw_f16x2(const __grid_constant__ Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int pix = a.Hi * a.I * a.Wi;          // 16-byte chunks per channel-chunk plane
  const int CCS = pix * 16;                   // bytes between channel chunks
  const int PLANE = a.cpt * CCS;              // bytes of one variant
  uint8_t* planes = smem;                     // [2 sets][4][cpt][Hi][I][Wi][16 B]: x hi, x lo, x^2 hi, x^2 lo of THIS CTA's tile
  uint8_t* Bs = planes + 2 * 4 * PLANE;       // [SB2][B_STAGE2]: this CTA's half of the weight rows
  float* bias_s = (float*)(Bs + SB2 * B_STAGE2);  // [2][BN]
  float* wsc_s = bias_s + 2 * BN;                // [2][BN] inverse weight-column scales
  float* xsc = wsc_s + 2 * BN;                   // [4 passes][16] inverse input scale per image
  float* xsf = xsc + 64;                         // [16] forward input scale per image (producers)
  unsigned* xmax = (unsigned*)(xsf + 16);        // [16] max |x| per image (float bits)
  uint64_t* bars = (uint64_t*)(((uintptr_t)(xmax + 16) + 7) & ~(uintptr_t)7);
  uint64_t* pfull = bars;                 // [2 sets] leader: all four planes of a set ready in BOTH CTAs
  uint64_t* pempty = pfull + 2;           // [2 sets] (multicast commit)
  uint64_t* bfull = pempty + 2;           // [SB2] local half stage landed
  uint64_t* bempty = bfull + SB2;         // [SB2] (multicast commit)
  uint64_t* tfull = bempty + SB2;         // [2] both accumulators of a tile complete (multicast commit)
  uint64_t* tempty = tfull + 2;           // [2] leader: drained by the epilogues of BOTH CTAs
  uint64_t* bfull_peer = tempty + 2;      // [SB2] leader: the peer's bfull, forwarded
  uint32_t* tmem_slot = (uint32_t*)(bfull_peer + SB2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const uint32_t rank = cluster_ctarank();
  const int HoWo = a.Ho * 8;
  const int64_t ntiles = (a.B + a.I - 1) / a.I;
  const int64_t npass = (ntiles + 1) / 2;         // a pass = the pair's two tiles
  const int64_t cl0 = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int nstages = (a.steps + G2 - 1) / G2;

  if (threadIdx.x == 0) {
    for (int p = 0; p < 2; ++p) { mbar_init(&pfull[p], 2 * PROD_WARPS2); mbar_init(&pempty[p], 1); }
    for (int s = 0; s < SB2; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); mbar_init(&bfull_peer[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 2 * EPI_WARPS2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS2)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int n = threadIdx.x; n < BN; n += THREADS2) {
    bias_s[n] = n < a.N ? a.bm[n] : 0.f;
    bias_s[BN + n] = n < a.N ? a.vb[n] : 1.f;
    wsc_s[n] = a.wsc[n];
    wsc_s[BN + n] = a.wsc[BN + n];
  }
  if (threadIdx.x < 16) xmax[threadIdx.x] = 0u;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before anyone arrives remotely / commits by multicast
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler

  if (warp < 4 || (warp >= 8 && warp < 12)) {
    // ================================================================== epilogue warps (fused ReLU + 2x2 pool)
    const int quad = warp & 3, chalf = warp >= 8 ? 1 : 0;
    const int r = quad * 32 + lane;             // TMEM lane = tile row (ho, img, wo)
    const int wo = r & 7, g = r >> 3, img = g % a.I, ho = g / a.I;
    const int hbit = 8 * a.I;
    const bool writer = (lane & 1) == 0 && (lane & hbit) == 0;
    const int Hp = a.Ho >> 1;
    int it = 0;
    for (int64_t ps = cl0; ps < npass; ps += ncl, ++it) {
      const int64_t braw = (ps * 2 + rank) * a.I + img;
      const bool st = braw < a.B && writer;
      const int64_t brow = braw < a.B ? braw : a.B - 1;  // shuffles need every lane: compute on a valid row
      const int buf = it & 1;
      const uint32_t outer = (uint32_t)(ho * 8 + wo);
      const int64_t m = brow * HoWo + outer;
      const uint64_t grow = a.eps.row_offset + (uint64_t)brow;
      const float* zrow = a.z ? a.z + brow * a.N : nullptr;
      float* dst = a.y + ((brow * Hp + (ho >> 1)) * 4 + (wo >> 1)) * a.N;
      mbar_wait_cl(&tfull[buf], (it >> 1) & 1);
      tc_fence_after();
      const float sx = xsc[(it & 3) * 16 + img], sx2 = sx * sx;
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 2 * BN);
      bool released = false;
#pragma unroll 1
      for (int c0 = chalf * 32; c0 < chalf * 32 + 32; c0 += 16) {
        if (c0 >= a.N) break;
        float mv[16], vv[16];
        tmem_ld16(trow + c0, mv);
        tmem_ld16(trow + BN + c0, vv);
        if (c0 + 16 >= a.N || c0 + 16 >= chalf * 32 + 32) {  // last chunk of this warp: accumulators are in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cl(&tempty[buf], 0);
          released = true;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int nb = c0 + 4 * q;
          if (nb >= a.N) break;
          float e[4];
          if (a.eps.injected) {
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = nb + j < a.N ? a.eps.injected[m * a.N + nb + j] : 0.f;
          } else {
            philox_normal4(a.eps, grow, outer, (uint32_t)(nb >> 2), e);
          }
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = nb + j < a.N ? nb + j : a.N - 1;
            float mean = mv[4 * q + j] * (sx * wsc_s[n]) + bias_s[n];
            const float sd = sqrt_fast(vv[4 * q + j] * (sx2 * wsc_s[BN + n]) + bias_s[BN + n]);
            if (zrow) mean *= __ldg(zrow + n);
            float v = fmaf(sd, e[j], mean);
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));     // wo pair
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, hbit));  // ho pair
            o[j] = fmaxf(v, 0.f);
          }
          if (st) {
            *reinterpret_cast<float2*>(dst + nb) = make_float2(o[0], o[1]);
            if (nb + 2 < a.N) *reinterpret_cast<float2*>(dst + nb + 2) = make_float2(o[2], o[3]);
          }
        }
      }
      if (!released) {  // a column half with nothing in it (N <= 32): still one arrival per warp and tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cl(&tempty[buf], 0);
      }
    }
  } else if (warp == 4) {
    // =============================================== MMA issuer (leader) / weight-stage forwarding (peer)
    if (rank != 0) {
      int st = 0;
      uint32_t bph = 0;
      for (int64_t ps = cl0; ps < npass; ps += ncl)
        for (int k = 0; k < 2 * nstages; ++k) {
          mbar_wait_cl(&bfull[st], bph);
          if (lane == 0) mbar_arrive_cl(&bfull_peer[st], 0);
          __syncwarp();
          if (++st == SB2) { st = 0; bph ^= 1; }
        }
    } else {
      // instruction descriptor: D = F32 (bit 4), A = B = F16 (format 0), K-major, N >> 3, M >> 4 with M = 256 over the pair
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint64_t bdesc0 = make_desc(smem_u32(Bs), BN * 16, 128);  // kc stride: 32 hi + 32 lo rows of 16 B
      const uint32_t plane0 = smem_u32(planes);
      int it = 0, st = 0;
      uint32_t bph = 0;
      long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long t_begin = clock64();
#define TIMED(k, stmt) { const long long c0_ = clock64(); stmt; tw[k] += clock64() - c0_; }
      for (int64_t ps = cl0; ps < npass; ps += ncl, ++it) {
        const int buf = it & 1, set = it & 1;
        for (int ph = 0; ph < 2; ++ph) {
          for (int sg = 0; sg < nstages; ++sg) {
            TIMED(0, mbar_wait_cl(&bfull[st], bph));
            TIMED(1, mbar_wait_cl(&bfull_peer[st], bph));
            if (sg == 0 && ph == 0) {
              TIMED(2, mbar_wait_cl(&tempty[buf], ((it >> 1) & 1) ^ 1));  // accumulators drained by both epilogues
              TIMED(4, mbar_wait_cl(&pfull[set], (it >> 1) & 1));          // planes of both tiles filled
            }
            tc_fence_after();
            const uint32_t dacc = tmem_base + (uint32_t)(buf * 2 * BN + ph * BN);
            const uint64_t abase = (uint64_t)((plane0 + (uint32_t)((set * 4 + 2 * ph) * PLANE)) >> 4);
            const uint64_t bst = bdesc0 + (uint64_t)((st * B_STAGE2) >> 4);
            const int s0 = sg * G2, ns = min(G2, a.steps - s0);
#pragma unroll
            for (int j = 0; j < G2; ++j) {
              if (j < ns) {
                const uint64_t ah = a.dtab[s0 + j] + abase, al = ah + (uint64_t)(PLANE >> 4);
                const uint64_t bh = bst + (uint64_t)((j * STEP_B2) >> 4);
                umma2_f16_elect(dacc, ah, bh, idesc, (s0 + j) == 0 ? 0u : 1u);       // x_hi * W_hi
                umma2_f16_elect(dacc, al, bh, idesc, 1u);                            // x_lo * W_hi
                umma2_f16_elect(dacc, ah, bh + (uint64_t)(((BN / 2) * 16) >> 4), idesc, 1u);  // x_hi * W_lo
              }
            }
            if (sg == nstages - 1 && ph == 1) {
              umma2_commit_mc_elect(&pempty[set]);  // the set's planes may be refilled (both CTAs)
              umma2_commit_mc_elect(&tfull[buf]);   // both accumulators of both tiles complete
            }
            umma2_commit_mc_elect(&bempty[st]);
            __syncwarp();
            if (++st == SB2) { st = 0; bph ^= 1; }
          }
        }
      }
#undef TIMED
      if (blockIdx.x == 0 && lane == 0) {
        tw[7] = clock64() - t_begin;
        for (int k = 0; k < 8; ++k) g_pair_prof[k] = (unsigned long long)tw[k];
      }
    }
  } else if (warp == 12) {
    // ==================================================================== weight loader (this CTA's half)
    if (lane == 0) {
      int st = 0;
      uint32_t bph = 0;
      const float* rbase = a.Bp + (size_t)rank * 2 * a.steps * (STEP_B2 / 4);
      for (int64_t ps = cl0; ps < npass; ps += ncl) {
        for (int ph = 0; ph < 2; ++ph) {
          const float* src0 = rbase + (size_t)ph * a.steps * (STEP_B2 / 4);
          for (int sg = 0; sg < nstages; ++sg) {
            const int ns = min(G2, a.steps - sg * G2);
            mbar_wait_cl(&bempty[st], bph ^ 1);
            mbar_arrive_expect_tx(&bfull[st], (uint32_t)(ns * STEP_B2));
            bulk_g2s(Bs + st * B_STAGE2, src0 + (size_t)sg * G2 * (STEP_B2 / 4), (uint32_t)(ns * STEP_B2), &bfull[st]);
            if (++st == SB2) { st = 0; bph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================== plane producers (this CTA's tile)
    // One sweep per pass: the tile is loaded ONCE and all four planes (x hi / lo, x^2 hi / lo) of the set are
    // written; sets alternate, so a tile is staged a full pass ahead of its MMAs.
    const int pt = (warp < 8 ? warp - 5 : warp - 10) * 32 + lane;  // 0 .. 383
    constexpr int NP = PROD_WARPS2 * 32;
    const int nchunks = a.cpt * pix;       // 16-byte chunks per variant (<= MAXC * NP, host-checked)
    const int img_f = a.H * a.W * a.C;     // floats per input image
    int o[MAXC], gsrc[MAXC], imx[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int q = pt + c * NP;
      o[c] = -1; gsrc[c] = -1; imx[c] = 0;
      if (q < nchunks) {
        const int cc = q % a.cpt;
        int p = q / a.cpt;
        const int wi = p % a.Wi;
        p /= a.Wi;
        const int im = p % a.I, hi = p / a.I;
        o[c] = cc * CCS + ((hi * a.I + im) * a.Wi + wi) * 16;
        imx[c] = im | (cc * 8 + 4 < a.C ? 256 : 0);
        if (hi < a.H && wi < a.W) gsrc[c] = (hi * a.W + wi) * a.C + cc * 8;
        if (a.half && cc == a.cpt - 1 && hi < a.H && wi + 1 < a.W) imx[c] |= 512;  // second half: next pixel
      }
    }
    int it = 0;
    for (int64_t ps = cl0; ps < npass; ps += ncl, ++it) {
      const int64_t b0 = (ps * 2 + rank) * a.I;
      const int set = it & 1;
      float4 v[MAXC][2];
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        v[c][0] = v[c][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gsrc[c] >= 0) {
          int64_t b = b0 + (imx[c] & 255);
          if (b >= a.B) b = a.B - 1;
          const float* src = a.x + b * img_f + gsrc[c];
          v[c][0] = __ldg(reinterpret_cast<const float4*>(src));
          if (imx[c] & 256) v[c][1] = __ldg(reinterpret_cast<const float4*>(src + 4));
          else if (imx[c] & 512) v[c][1] = __ldg(reinterpret_cast<const float4*>(src + a.C));
        }
      }
      // ---- max |x| per image -> power-of-two scales
      for (int im = 0; im < a.I; ++im) {
        float m = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (o[c] >= 0 && (imx[c] & 255) == im) {
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v[c][0].x), fabsf(v[c][0].y)), fmaxf(fabsf(v[c][0].z), fabsf(v[c][0].w))));
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v[c][1].x), fabsf(v[c][1].y)), fmaxf(fabsf(v[c][1].z), fabsf(v[c][1].w))));
          }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sh));
        if (lane == 0) atomicMax(xmax + im, __float_as_uint(m));
      }
      asm volatile("bar.sync 2, %0;" ::"n"(NP) : "memory");
      if (pt < a.I) {
        const float m = __uint_as_float(xmax[pt]);
        int e = 0;
        if (m > 0.f && m < 3.0e38f) {
          int ex;
          frexpf(m, &ex);          // m = f * 2^ex, f in [0.5, 1)
          e = 7 - ex;              // m * 2^e in [2^6, 2^7)
          e = e > 100 ? 100 : (e < -100 ? -100 : e);
        }
        xsf[pt] = ldexpf(1.f, e);
        xsc[(it & 3) * 16 + pt] = ldexpf(1.f, -e);
        xmax[pt] = 0u;  // ready for the next pass
      }
      asm volatile("bar.sync 2, %0;" ::"n"(NP) : "memory");
      mbar_wait_cl(&pempty[set], ((it >> 1) & 1) ^ 1);
      uint8_t* base = planes + (size_t)set * 4 * PLANE;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (o[c] < 0) continue;
        const float f = xsf[imx[c] & 255];
        const float in[8] = {v[c][0].x, v[c][0].y, v[c][0].z, v[c][0].w, v[c][1].x, v[c][1].y, v[c][1].z, v[c][1].w};
        uint32_t hw[4], lw[4], hs[4], ls[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a0 = in[2 * e] * f, a1 = in[2 * e + 1] * f;
          split_f16x2(a0, a1, hw[e], lw[e]);
          split_f16x2(a0 * a0, a1 * a1, hs[e], ls[e]);
        }
        *reinterpret_cast<uint4*>(base + o[c]) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(base + PLANE + o[c]) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        *reinterpret_cast<uint4*>(base + 2 * PLANE + o[c]) = make_uint4(hs[0], hs[1], hs[2], hs[3]);
        *reinterpret_cast<uint4*>(base + 3 * PLANE + o[c]) = make_uint4(ls[0], ls[1], ls[2], ls[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cl(&pfull[set], 0);
      // the scales of this pass are consumed (xsf) before the next pass overwrites them
      asm volatile("bar.sync 2, %0;" ::"n"(NP) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still touch its shared / tensor memory
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS2) : "memory");
  }
}

// weights -> Bp2[rank][phase][step][kc 2][hi n 32 | lo n 32][8 k] fp16: CTA `rank` of a pair holds columns [32 rank, 32 rank + 32)
__global__ void pack_f16_weights_x2(int N, int steps, const int* __restrict__ kmap, const float* __restrict__ Wm,
                                    const float* __restrict__ Ls, float max_std, const int* __restrict__ esc,
                                    __half* __restrict__ Bp) {
  const int64_t total = (int64_t)2 * steps * 2 * BN * 8;  // (phase, step, kc, n, e)
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx & 7);
    int64_t t = idx >> 3;
    const int n = (int)(t % BN);
    t /= BN;
    const int kc = (int)(t & 1);
    t >>= 1;
    const int s = (int)(t % steps), ph = (int)(t / steps);
    const int k = kmap[s * 16 + kc * 8 + e];
    float w = 0.f;
    if (k >= 0 && n < N) {
      if (ph == 0) w = Wm[(int64_t)k * N + n];
      else {
        const float sd = fminf(expf(Ls[(int64_t)k * N + n]), max_std);
        w = sd * sd;
      }
      w = ldexpf(w, esc[ph * BN + n]);
    }
    const __half hi = __float2half_rn(w), lo = __float2half_rn(w - __half2float(hi));
    const int rk = n / (BN / 2), nl = n % (BN / 2);
    __half* base = Bp + (((int64_t)rk * 2 + ph) * steps + s) * (STEP_B2 / 2) + ((int64_t)kc * BN + nl) * 8 + e;
    base[0] = hi;
    base[(BN / 2) * 8] = lo;  // the 32 lo rows follow the 32 hi rows of the same K chunk
  }
}

// column maxima -> power-of-two scales.  esc[ph*BN + n] = e with max_k |w[k,n]| * 2^e in [2^13, 2^14);
// wsc = 2^-e (the epilogue's inverse scale); also the clipped bias variance vb.
__global__ void f16_colscale(int K, int N, const float* __restrict__ Wm, const float* __restrict__ Ls, float max_std,
                             int* __restrict__ esc, float* __restrict__ wsc, const float* __restrict__ lvb,
                             float* __restrict__ vb) {
  const int n = blockIdx.x, ph = blockIdx.y;
  float m = 0.f;
  if (n < N)
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float w;
      if (ph == 0) w = fabsf(Wm[(int64_t)k * N + n]);
      else {
        const float sd = fminf(expf(Ls[(int64_t)k * N + n]), max_std);
        w = sd * sd;
      }
      m = fmaxf(m, w);
    }
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    int e = 0;
    if (m > 0.f && m < 3.0e38f) {
      int ex;
      frexpf(m, &ex);
      e = 14 - ex;
      e = e > 100 ? 100 : (e < -100 ? -100 : e);
    }
    esc[ph * BN + n] = e;
    wsc[ph * BN + n] = ldexpf(1.f, -e);
    if (ph == 0 && n < N) vb[n] = fminf(expf(lvb[n]), max_std * max_std);
  }
}

// weights -> Bp[phase][step][kc 2][hi n 64 | lo n 64][8 k] fp16; kmap[step][16] = global k index or -1
__global__ void pack_f16_weights(int N, int steps, const int* __restrict__ kmap, const float* __restrict__ Wm,
                                 const float* __restrict__ Ls, float max_std, const int* __restrict__ esc,
                                 __half* __restrict__ Bp) {
  const int64_t total = (int64_t)2 * steps * 2 * BN * 8;  // (phase, step, kc, n, e)
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx & 7);
    int64_t t = idx >> 3;
    const int n = (int)(t % BN);
    t /= BN;
    const int kc = (int)(t & 1);
    t >>= 1;
    const int s = (int)(t % steps), ph = (int)(t / steps);
    const int k = kmap[s * 16 + kc * 8 + e];
    float w = 0.f;
    if (k >= 0 && n < N) {
      if (ph == 0) w = Wm[(int64_t)k * N + n];
      else {
        const float sd = fminf(expf(Ls[(int64_t)k * N + n]), max_std);
        w = sd * sd;
      }
      w = ldexpf(w, esc[ph * BN + n]);
    }
    const __half hi = __float2half_rn(w), lo = __float2half_rn(w - __half2float(hi));
    __half* base = Bp + ((int64_t)ph * steps + s) * (STEP_B / 2) + ((int64_t)kc * 2 * BN + n) * 8 + e;
    base[0] = hi;
    base[BN * 8] = lo;  // the lo rows follow the 64 hi rows of the same K chunk
  }
}

}  // namespace cf

// Returns MNF_OK with *used = false when the shape is outside this kernel's domain.
int mnf_conv_raw_f16_forward(mnf_ctx* ctx, int64_t B, const int* geo, int N, const float* x, const float* z,
                              const float* mean_W, const float* log_std_W, const float* mean_b,
                              const float* log_var_b, float max_std, const NoiseDev& eps, float* y, float* mean_out,
                              float* sd_out, bool* used) {
  using namespace cf;
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
  a.cpt = (C + 7) / 8;
  a.pool = pool;
  // ---- step table: chunk pairs of one tap, then the odd leftover chunks of neighbouring taps
  const int pixn = a.Hi * a.I * a.Wi, CCS = pixn * 16;
  std::vector<int> tab, kmap;
  auto chunk_off = [&](int i, int j, int cc) { return cc * CCS + (i * a.I * a.Wi + j) * 16; };
  auto push_k = [&](int i, int j, int cc) {
    for (int e = 0; e < 8; ++e) kmap.push_back(i >= 0 && cc * 8 + e < C ? (i * kw + j) * C + cc * 8 + e : -1);
  };
  // C % 8 == 4 (MNF-LeNet conv2: C = 20): the last chunk of a pixel is only half used.  Its plane
  // then holds [x[p][C-4..C) | x[p+1][C-4..C)] (p+1: the next pixel of the row), so ONE chunk serves
  // the leftover channels of two horizontally adjacent taps and a filter row needs (kw+1)/2 of them.
  a.half = ((C & 7) == 4 && (a.cpt & 1)) ? 1 : 0;
  auto push_k_half = [&](int i, int j, int cc) {
    for (int e = 0; e < 4; ++e) kmap.push_back((i * kw + j) * C + cc * 8 + e);
    for (int e = 0; e < 4; ++e) kmap.push_back(j + 1 < kw ? (i * kw + j + 1) * C + cc * 8 + e : -1);
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
      if ((a.cpt & 1) && (!a.half || (j & 1) == 0)) left.push_back({i, j});
    }
  auto push_left = [&](const std::pair<int, int>& t) {
    if (a.half) push_k_half(t.first, t.second, a.cpt - 1);
    else push_k(t.first, t.second, a.cpt - 1);
  };
  for (size_t m = 0; m < left.size(); m += 2) {
    const int cc = a.cpt - 1;
    const int o0 = chunk_off(left[m].first, left[m].second, cc);
    tab.push_back(o0);
    push_left(left[m]);
    if (m + 1 < left.size()) {
      const int o1 = chunk_off(left[m + 1].first, left[m + 1].second, cc);
      if (o1 <= o0 || (o1 - o0) >= (1 << 18)) return MNF_OK;
      tab.push_back(o1 - o0);
      push_left(left[m + 1]);
    } else {
      tab.push_back(16);   // dummy second chunk: valid memory, zero weights
      push_k(-1, 0, 0);
    }
  }
  a.steps = (int)tab.size() / 2;
  if (a.steps > MAXSTEPS) return MNF_OK;
  for (int i = 0; i < a.steps; ++i)  // K-major SWIZZLE_NONE descriptor (umma.cuh:make_desc) with start = step offset
    a.dtab[i] = (uint64_t)(((uint32_t)tab[2 * i] >> 4) & 0x3FFF) | ((uint64_t)(((uint32_t)tab[2 * i + 1] >> 4) & 0x3FFF) << 16) |
                ((uint64_t)(((uint32_t)(a.Wi * 16) >> 4) & 0x3FFF) << 32) | (1ull << 46);
  if (a.cpt * pixn > MAXC * PROD_WARPS * 32) return MNF_OK;  // the producers hold a whole tile in registers
  const size_t plane_bytes = (size_t)TT * 4 * a.cpt * CCS;
  const size_t smem = plane_bytes + (size_t)SB * B_STAGE + sizeof(float) * BM * BN + sizeof(int) * 2 * a.steps +
                      sizeof(float) * ((2 + 16 + 2) * BN + 128 + 32 + 32) + 256;
  if (smem > 227 * 1024) return MNF_OK;
  // CTA-pair kernel (cta_group::2) for the fused-pool inference shape with enough tiles to feed every pair
  static const bool no_pair = getenv("MNF_CONV_NO_PAIR") != nullptr;
  const bool pair = !no_pair && pool && a.I <= 2 && !sd_out && !mean_out && !(N & 1) && cdiv(B, a.I) >= 2 * (int64_t)ctx->sm_count;
  // ---- workspace: Bp | vb | wsc | esc | steptab | kmap
  const size_t bp_floats = (size_t)2 * a.steps * (STEP_B / 4);  // same size in both layouts
  const size_t ws_bytes = sizeof(float) * (bp_floats + 64 + 2 * BN + 2 * BN) + sizeof(int) * (tab.size() + kmap.size() + 128);
  bool fresh = true;
  char* ws = (char*)mnf_packed(ctx, mnf_mix(pair ? 0xC0F162ull : 0xC0F16ull, mnf_mix(((uint64_t)H << 40) | ((uint64_t)W << 20) | (uint32_t)C,
                                                               ((uint64_t)kh << 40) | ((uint64_t)kw << 20) | (uint32_t)N)),
                               (uint64_t)(uintptr_t)mean_W, (uint64_t)(uintptr_t)log_std_W,
                               mnf_mix((uint64_t)(uintptr_t)log_var_b, (uint64_t)__float_as_uint_host(max_std)), ws_bytes, &fresh);
  if (!ws) return MNF_ERR_CUDA;
  float* Bp = (float*)ws;
  float* vb = Bp + bp_floats;
  float* wsc = vb + 64;
  int* esc = (int*)(wsc + 2 * BN);
  int* d_tab = esc + 2 * BN;
  int* d_kmap = d_tab + ((tab.size() + 63) & ~(size_t)63);
  if (fresh) {
    MNF_CUDA(cudaMemcpyAsync(d_tab, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice, ctx->stream));
    MNF_CUDA(cudaMemcpyAsync(d_kmap, kmap.data(), sizeof(int) * kmap.size(), cudaMemcpyHostToDevice, ctx->stream));
    f16_colscale<<<dim3(BN, 2), 128, 0, ctx->stream>>>(kh * kw * C, N, mean_W, log_std_W, max_std, esc, wsc, log_var_b, vb);
    MNF_LAUNCHED(ctx);
    const int64_t tot = (int64_t)2 * a.steps * 2 * BN * 2;
    int grid = (int)(cdiv(tot, 256) < ctx->sm_count * 4 ? cdiv(tot, 256) : ctx->sm_count * 4);
    if (pair) pack_f16_weights_x2<<<grid, 256, 0, ctx->stream>>>(N, a.steps, d_kmap, mean_W, log_std_W, max_std, esc, (__half*)Bp);
    else pack_f16_weights<<<grid, 256, 0, ctx->stream>>>(N, a.steps, d_kmap, mean_W, log_std_W, max_std, esc, (__half*)Bp);
    MNF_LAUNCHED(ctx);
  }
  a.x = x; a.z = z; a.bm = mean_b; a.vb = vb; a.steptab = d_tab; a.Bp = Bp; a.wsc = wsc;
  a.y = y; a.mean_out = mean_out; a.sd_out = sd_out; a.eps = eps;
  if (pair) {
    const size_t smem2 = (size_t)2 * 4 * a.cpt * CCS + (size_t)SB2 * B_STAGE2 + sizeof(float) * (4 * BN + 64 + 16 + 16) + 512;
    static size_t attr_smem2 = 0;
    if (smem2 > attr_smem2) {
      MNF_CUDA(cudaFuncSetAttribute(conv_raw_f16x2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      attr_smem2 = smem2;
    }
    const int64_t passes2 = cdiv(cdiv(B, a.I), (int64_t)2);
    const int64_t pairs = ctx->sm_count / 2;
    const int g2 = 2 * (int)(passes2 < pairs ? passes2 : pairs);
    const bool probed = mnf_probe_begin(ctx, MNF_PROBE_CONV_GEMM);
    mnf_note_path("conv_raw_f16x2");
    conv_raw_f16x2<<<g2, THREADS2, smem2, ctx->stream>>>(a);
    mnf_probe_end(ctx, probed);
    MNF_LAUNCHED(ctx);
    static const bool prof = getenv("MNF_CONV_PROF") != nullptr;
    if (prof) {
      unsigned long long h[8];
      cudaStreamSynchronize(ctx->stream);
      cudaMemcpyFromSymbol(h, g_pair_prof, sizeof(h));
      fprintf(stderr, "[conv_raw_f16x2 leader MMA warp, cluster 0] wait cycles: bfull %llu bfull_peer %llu tempty %llu "
                      "pfull %llu | total %llu (passes %lld)\n", h[0], h[1], h[2], h[4], h[7], (long long)cdiv(passes2, pairs));
    }
    *used = true;
    return MNF_OK;
  }
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MNF_CUDA(cudaFuncSetAttribute(conv_raw_f16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int64_t passes = cdiv(cdiv(B, a.I), (int64_t)TT);
  const int g = (int)(passes < ctx->sm_count ? passes : ctx->sm_count);
  const bool probed = mnf_probe_begin(ctx, MNF_PROBE_CONV_GEMM);
  mnf_note_path("conv_raw_f16");
  conv_raw_f16<<<g, THREADS, smem, ctx->stream>>>(a);
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  *used = true;
  return MNF_OK;
}
