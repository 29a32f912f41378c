// RNVP flow step (rnvp.py:33-47) on tcgen05 kind::f16 for many rows: the coupling MLP of one step for a
// 128-row tile in ONE CTA, every operand split x = hi + lo into fp16 planes (hi*hi + lo*hi + hi*lo, fp32
// accumulation in TMEM: per-term error ~2^-22), z moved between HBM and shared memory ONLY by TMA tensor copies
// (cp.async.bulk.tensor.2d, one instruction per [128 rows x 32 columns] tile, SWIZZLE_128B so that a thread
// reading its own row is conflict-free; out-of-range rows / columns are zero-filled on load and clipped on
// store by the hardware):
//
//   loader warp      z[128 x 32] tiles -> raw ring (6 x 16 KB), pre-split weight blocks -> operand rings;
//                    stage 1 and stage 2 walk the same ring
//   mask warps 0-3   the step's Bernoulli(0.5) mask bits of the tile, once (one Philox call per 128 elements)
//   stage 1          16 converter warps (thread = row): raw -> mask -> * 2^e -> fp16 hi/lo -> K-major A tiles;
//                    acc_h[128 x 64] = (m * z) W1 by the MMA warp
//   stage 1e         16 warps: h = tanh(acc_h / scales + b1) -> * 2^13 -> fp16 hi/lo A operand of stage 2
//   stage 2          per 32 columns: acc[128 x (32 s | 32 t)] = h [Ws | Wt], double-buffered in TMEM;
//                    16 epilogue warps read z from the raw tile, apply the gated update IN PLACE, and the
//                    tile leaves by one TMA tensor store (no row-strided global access: the tf32 kernel this
//                    replaces wrote 98 MB for 33 MB of output and stalled on its per-thread global loads;
//                    a first version with one 1-D bulk copy per ROW was issue-bound on the copies -- UBLKCP
//                    takes uniform-register operands, so per-lane addresses serialise: 8 k cycles per tile)
//
// fp16 lacks range: z is scaled by a power of two taken from the tile's max |z| (produced by whoever wrote z:
// sample_z0_tmax, the previous step's epilogue, or tile_absmax), each weight matrix by its own maximum, h (in
// [-1, 1]) by 2^13; the epilogues multiply by the exact inverses.  flows_umma.cu (3xTF32) keeps IAF / MADE steps
// and shapes this kernel does not take (D % 4 != 0).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "umma.cuh"

namespace f2 {
using namespace um;

constexpr int BM = 128, HP = 64, TK = 32;
constexpr int RAW_TILE = BM * TK * 4;  // 16 KB: [128 rows][128 B], 16-byte chunk q of row r at chunk q ^ (r & 7)
constexpr int NR = 6, S1 = 4, S2 = 3, NG = 4;  // NR raw tiles in flight: four being converted + two on their way (TMA latency)
constexpr int A1_VAR = (TK / 8) * BM * 16, A1_STAGE = 2 * A1_VAR;  // 8 KB / 16 KB
constexpr int B1_VAR = (TK / 8) * HP * 16, B1_STAGE = 2 * B1_VAR;  // 4 KB / 8 KB
constexpr int H_VAR = (HP / 8) * BM * 16;                          // 16 KB
constexpr int B2_VAR = (HP / 8) * 64 * 16, B2_STAGE = 2 * B2_VAR;  // 8 KB / 16 KB
constexpr int REGION_X = S1 * (A1_STAGE + B1_STAGE);               // 96 KB: stage-1 rings, then B2 ring + h
static_assert(S2 * B2_STAGE + 2 * H_VAR <= REGION_X, "stage-2 buffers alias the stage-1 rings");
static_assert(S1 == NG && NR >= NG, "group g owns operand stage g; every group can hold a raw tile");
constexpr int WARPS = 22, THREADS = WARPS * 32;  // 0-3 masks + epilogues, 4 MMA, 5-20 converters + epilogues, 21 loader
constexpr int TMEM_COLS = 256;                   // acc_h 64 + 2 x 64
constexpr float H_SCALE = 8192.f, H_INV = 1.f / 8192.f;

struct Args {
  int64_t B;
  int D, H, kblocks;  // kblocks = ceil(D / 32): stage-1 k-blocks == stage-2 chunks
  const float* zin;
  float* zout;
  const float* ld_in;
  float* ld_out;
  float* hid_save;
  const float* b1;
  const float* bs;
  const float* bt;
  const uint8_t* W1p;   // [kblocks][{hi, lo}][4 chunks][64 hidden][8 halves]
  const uint8_t* W2p;   // [kblocks][{hi, lo}][8 chunks][32 s | 32 t][8 halves]
  const float* sinv;    // [3] inverse power-of-two scales of W1, Ws, Wt
  const float* tmax_in; // [tiles] max |z| of each 128-row tile
  float* tmax_out;      // [tiles] same for zout (may be NULL)
  NoiseDev bern;
};

__device__ __forceinline__ void split_f16x2(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a0, a1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 2^e with m * 2^e in [2^13, 2^14) (e = 0 for m == 0 / non-finite)
__host__ __device__ __forceinline__ int scale_exp(float m) {
  int e = 0;
  if (m > 0.f && m < 3.0e38f) {
    int ex;
    frexpf(m, &ex);
    e = 14 - ex;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
  }
  return e;
}
__device__ __forceinline__ void tma_load_2d(void* sdst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(sdst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, const void* ssrc) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1),
               "r"(smem_u32(ssrc))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
// single MUFU instructions (no denormal / range fix-ups: the arguments here are normal and bounded)
__device__ __forceinline__ float ex2_fast(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__global__ void __launch_bounds__(THREADS, 1) rnvp_f16_step(const __grid_constant__ Args a, const __grid_constant__ CUtensorMap tm_in,
                                                             const __grid_constant__ CUtensorMap tm_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* A1 = smem;                        // [S1][2][4][BM][16 B]   64 KB
  uint8_t* B1 = smem + S1 * A1_STAGE;        // [S1][2][4][HP][16 B]   32 KB
  uint8_t* B2 = smem;                        // [S2][2][8][64][16 B]   48 KB (aliases A1)
  uint8_t* Hs = smem + S2 * B2_STAGE;        // [2][8][BM][16 B]       32 KB (aliases A1 / B1)
  uint8_t* raw = smem + REGION_X;            // [NR][BM][128 B] (swizzled)
  uint32_t* mbits = (uint32_t*)(raw + NR * RAW_TILE);  // [kblocks][BM]: bit e set = element kept (m = 1)
  float* colp = (float*)(mbits + a.kblocks * BM);      // [2][kblocks * 32]: bs | bt
  float* b1s = colp + 2 * a.kblocks * TK;              // [64] b1
  float* ldp = b1s + HP;                               // [4][BM] log-det partials
  float* zmx = ldp + 4 * BM;                           // [32] per-warp max |z'|
  uint64_t* bars = (uint64_t*)(zmx + 32);
  uint64_t* full1 = bars;              // [S1]
  uint64_t* empty1 = full1 + S1;       // [S1]
  uint64_t* raw_full = empty1 + S1;    // [NR]
  uint64_t* raw_free = raw_full + NR;  // [NR]
  uint64_t* acc1_full = raw_free + NR; // [1]
  uint64_t* h_ready = acc1_full + 1;   // [1]
  uint64_t* full2 = h_ready + 1;       // [S2]
  uint64_t* empty2 = full2 + S2;       // [S2]
  uint64_t* afull = empty2 + S2;       // [2]
  uint64_t* aempty = afull + 2;        // [2]
  uint32_t* tmem_slot = (uint32_t*)(aempty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int D = a.D, ncols = a.kblocks * TK;
  const int nrows = (int)(a.B - row0 < BM ? a.B - row0 : BM);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S1; ++s) { mbar_init(&full1[s], 4 + 1); mbar_init(&empty1[s], 1); }
    for (int s = 0; s < NR; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_free[s], 4); }
    mbar_init(acc1_full, 1);
    mbar_init(h_ready, 16);
    for (int s = 0; s < S2; ++s) { mbar_init(&full2[s], 1); mbar_init(&empty2[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&afull[b], 1); mbar_init(&aempty[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int c = threadIdx.x; c < ncols; c += THREADS) {
    colp[c] = c < D ? a.bs[c] : 0.f;
    colp[ncols + c] = c < D ? a.bt[c] : 0.f;
  }
  if (threadIdx.x < HP) b1s[threadIdx.x] = threadIdx.x < a.H ? a.b1[threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t t_acc1 = tmem_base, t_acc2 = tmem_base + 64;

  // power-of-two scale of this tile's z: max |z| * 2^e in [2^13, 2^14)
  const int ze = scale_exp(__ldg(a.tmax_in + blockIdx.x));
  const float zscale = ldexpf(1.f, ze);
  const float c1 = ldexpf(1.f, -ze) * __ldg(a.sinv), c2s = __ldg(a.sinv + 1) * H_INV, c2t = __ldg(a.sinv + 2) * H_INV;

  // instruction descriptor: D = F32 (bit 4), A = B = F16 (format 0), K-major both, N = 64, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  // ---- stage-1 epilogue: h = tanh(acc1 / scales + b1) -> 2^13 h as fp16 hi/lo, the A operand of stage 2
  auto h_epilogue = [&](int quad, int c0) {
    const int r = quad * 32 + lane;
    const int64_t row = row0 + r;
    mbar_wait(acc1_full, 0);
    tc_fence_after();
    float v[16];
    tmem_ld16(t_acc1 + ((uint32_t)(quad * 32) << 16) + c0, v);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float hv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int h = c0 + 8 * q + e;
        const float pre = fmaf(v[8 * q + e], c1, b1s[h]);
        const float t = ex2_fast(-2.8853900817779268f * fabsf(pre));  // exp(-2 |pre|) in (0, 1]
        const float th = (1.f - t) * rcp_fast(1.f + t);               // tanh(|pre|)
        hv[e] = h < a.H ? (pre < 0.f ? -th : th) : 0.f;
        if (a.hid_save && row < a.B && h < a.H) a.hid_save[row * a.H + h] = hv[e];
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_f16x2(hv[2 * e] * H_SCALE, hv[2 * e + 1] * H_SCALE, hi[e], lo[e]);
      const int kc = (c0 >> 3) + q;  // 16-byte chunk (8 hidden units) along K
      *reinterpret_cast<uint4*>(Hs + kc * (BM * 16) + r * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(Hs + H_VAR + kc * (BM * 16) + r * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(h_ready);
  };

  if (warp == WARPS - 1) {
    // =============================================================================== loader warp
    auto load_raw = [&](int seq, int colblk) {
      const int slot = seq % NR;
      mbar_wait(&raw_free[slot], (uint32_t)(((seq / NR) & 1) ^ 1));
      if (lane == 0) {
        mbar_arrive_expect_tx(&raw_full[slot], RAW_TILE);  // the whole box counts, zero-filled parts included
        tma_load_2d(raw + slot * RAW_TILE, &tm_in, colblk * TK, (int)row0, &raw_full[slot]);
      }
      __syncwarp();
    };
    for (int kb = 0; kb < a.kblocks; ++kb) {
      load_raw(kb, kb);
      const int stage = kb % S1;
      mbar_wait(&empty1[stage], (uint32_t)(((kb / S1) & 1) ^ 1));
      if (lane == 0) {
        mbar_arrive_expect_tx(&full1[stage], B1_STAGE);
        bulk_g2s(B1 + stage * B1_STAGE, a.W1p + (size_t)kb * B1_STAGE, B1_STAGE, &full1[stage]);
      }
      __syncwarp();
    }
    for (int c = 0; c < a.kblocks; ++c) {
      load_raw(a.kblocks + c, c);
      if (c == 0) mbar_wait(acc1_full, 0);  // the weight ring re-uses the stage-1 buffers
      const int s2 = c % S2;
      mbar_wait(&empty2[s2], (uint32_t)(((c / S2) & 1) ^ 1));
      if (lane == 0) {
        mbar_arrive_expect_tx(&full2[s2], B2_STAGE);
        bulk_g2s(B2 + s2 * B2_STAGE, a.W2p + (size_t)c * B2_STAGE, B2_STAGE, &full2[s2]);
      }
      __syncwarp();
    }
  } else if (warp >= 5) {
    // ============================================================ stage-1 converters (warps 5-20)
    const int pt = threadIdx.x - 5 * 32;  // 0..511: four groups of 128 rows
    const int r = pt & (BM - 1), grp = pt >> 7;
    named_sync(3, 128 + 512);  // mask bits are in shared memory
    for (int kb = grp; kb < a.kblocks; kb += NG) {
      const int slot = kb % NR;
      mbar_wait(&raw_full[slot], (uint32_t)((kb / NR) & 1));
      const uint8_t* src = raw + slot * RAW_TILE + r * 128;
      float v[TK];
#pragma unroll
      for (int q = 0; q < TK / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(src + ((q ^ (r & 7)) << 4));
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_free[slot]);
      const uint32_t bits = mbits[kb * BM + r];  // (columns beyond D: zero-filled by the load)
      uint32_t hi[TK / 2], lo[TK / 2];
#pragma unroll
      for (int e = 0; e < TK / 2; ++e) {
        const float x0 = (bits >> (2 * e)) & 1u ? v[2 * e] * zscale : 0.f;
        const float x1 = (bits >> (2 * e + 1)) & 1u ? v[2 * e + 1] * zscale : 0.f;
        split_f16x2(x0, x1, hi[e], lo[e]);
      }
      const int stage = kb % S1;
      mbar_wait(&empty1[stage], (uint32_t)(((kb / S1) & 1) ^ 1));
      uint8_t* dst = A1 + stage * A1_STAGE + r * 16;
#pragma unroll
      for (int kc = 0; kc < TK / 8; ++kc) {
        *reinterpret_cast<uint4*>(dst + kc * (BM * 16)) = make_uint4(hi[4 * kc], hi[4 * kc + 1], hi[4 * kc + 2], hi[4 * kc + 3]);
        *reinterpret_cast<uint4*>(dst + A1_VAR + kc * (BM * 16)) = make_uint4(lo[4 * kc], lo[4 * kc + 1], lo[4 * kc + 2], lo[4 * kc + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full1[stage]);
    }
    // ---- the converter warps that also serve stage 2 take their share of the stage-1 epilogue
    if (warp >= 8 && warp < 12) h_epilogue(warp & 3, 16);
    else if (warp >= 13 && warp < 17) h_epilogue(warp & 3, 32);
    else if (warp >= 17) h_epilogue(warp & 3, 48);
  } else if (warp == 4) {
    // ================================================================================ MMA warp
    const uint64_t a1d = make_desc(smem_u32(A1), BM * 16, 128), b1d = make_desc(smem_u32(B1), HP * 16, 128);
    int stage = 0;
    uint32_t sph = 0;
    for (int kb = 0; kb < a.kblocks; ++kb) {
      mbar_wait(&full1[stage], sph);
      tc_fence_after();
      const uint64_t ad = a1d + (uint64_t)((stage * A1_STAGE) >> 4), bd = b1d + (uint64_t)((stage * B1_STAGE) >> 4);
#pragma unroll
      for (int j = 0; j < TK / 16; ++j) {
        const uint64_t ah = ad + (uint64_t)((2 * j * BM * 16) >> 4), al = ah + (A1_VAR >> 4);
        const uint64_t bh = bd + (uint64_t)((2 * j * HP * 16) >> 4), bl = bh + (B1_VAR >> 4);
        umma_f16_elect(t_acc1, ah, bh, idesc, (kb == 0 && j == 0) ? 0u : 1u);
        umma_f16_elect(t_acc1, al, bh, idesc, 1u);
        umma_f16_elect(t_acc1, ah, bl, idesc, 1u);
      }
      umma_commit_elect(&empty1[stage]);
      if (kb == a.kblocks - 1) umma_commit_elect(acc1_full);
      __syncwarp();
      if (++stage == S1) { stage = 0; sph ^= 1; }
    }
    mbar_wait(h_ready, 0);
    tc_fence_after();
    const uint64_t hd = make_desc(smem_u32(Hs), BM * 16, 128), b2d = make_desc(smem_u32(B2), 64 * 16, 128);
    int s2 = 0;
    uint32_t p2 = 0;
    for (int c = 0; c < a.kblocks; ++c) {
      const int buf = c & 1;
      mbar_wait(&aempty[buf], (uint32_t)(((c >> 1) & 1) ^ 1));
      mbar_wait(&full2[s2], p2);
      tc_fence_after();
      const uint64_t bd = b2d + (uint64_t)((s2 * B2_STAGE) >> 4);
      const uint32_t dacc = t_acc2 + buf * 64;
#pragma unroll
      for (int j = 0; j < HP / 16; ++j) {
        const uint64_t ah = hd + (uint64_t)((2 * j * BM * 16) >> 4), al = ah + (H_VAR >> 4);
        const uint64_t bh = bd + (uint64_t)((2 * j * 64 * 16) >> 4), bl = bh + (B2_VAR >> 4);
        umma_f16_elect(dacc, ah, bh, idesc, j == 0 ? 0u : 1u);
        umma_f16_elect(dacc, al, bh, idesc, 1u);
        umma_f16_elect(dacc, ah, bl, idesc, 1u);
      }
      umma_commit_elect(&empty2[s2]);
      umma_commit_elect(&afull[buf]);
      __syncwarp();
      if (++s2 == S2) { s2 = 0; p2 ^= 1; }
    }
  } else {
    // ============================================== warps 0-3: the tile's mask bits, then the h epilogue
    const int r = warp * 32 + lane;
    const int64_t row = row0 + (r < nrows ? r : nrows - 1);
    for (int k4 = 0; k4 * 4 < a.kblocks; ++k4) {
      uint32_t w[4];
      if (a.bern.injected) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          w[i] = 0u;
          for (int e = 0; e < TK; ++e) {
            const int d = (k4 * 4 + i) * TK + e;
            if (d < D && a.bern.injected[row * D + d] != 0.f) w[i] |= 1u << e;
          }
        }
      } else {
        const Philox4 q = philox_at(a.bern, a.bern.row_offset + (uint64_t)row, 0u, (uint32_t)k4);
        w[0] = ~q.x; w[1] = ~q.y; w[2] = ~q.z; w[3] = ~q.w;  // the draw is 1 when the bit is clear
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kb = k4 * 4 + i, left = D - kb * TK;
        // columns beyond D count as kept (m = 1): z' = z = 0 (zero-filled by the TMA load) and no log-det term
        if (kb < a.kblocks) mbits[kb * BM + r] = left < TK ? (w[i] | ~((1u << left) - 1u)) : w[i];
      }
    }
    named_arrive(3, 128 + 512);
    h_epilogue(warp, 0);
  }

  // ================================= stage-2 epilogue: warps 0-3, 13-16 (columns 0-15), 8-11, 17-20 (16-31)
  float zm = 0.f;
  if (warp < 4 || (warp >= 8 && warp < 12) || (warp >= 13 && warp < WARPS - 1)) {
    const int g = warp & 3;
    const int half = (warp >= 8 && warp < 12) || warp >= 17 ? 1 : 0, par = warp >= 13 ? 1 : 0;
    const int r = g * 32 + lane;
    const bool row_ok = r < nrows;
    float ld = 0.f;
    for (int c = par; c < a.kblocks; c += 2) {
      const int buf = c & 1, seq = a.kblocks + c, slot = seq % NR;
      const int cb = c * TK + half * 16;
      uint8_t* zt = raw + slot * RAW_TILE + r * 128;
      const int sw = r & 7, q0 = half * 4;
      mbar_wait(&raw_full[slot], (uint32_t)((seq / NR) & 1));
      float z16[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(zt + (((q0 + q) ^ sw) << 4));
        z16[4 * q] = t.x; z16[4 * q + 1] = t.y; z16[4 * q + 2] = t.z; z16[4 * q + 3] = t.w;
      }
      const uint32_t mb = mbits[c * BM + r] >> (half * 16);
      mbar_wait(&afull[buf], (uint32_t)((c >> 1) & 1));
      tc_fence_after();
      const uint32_t tb = t_acc2 + buf * 64 + ((uint32_t)(g * 32) << 16) + half * 16;
      float sv[16], tv[16];
      tmem_ld16(tb, sv);
      tmem_ld16(tb + 32, tv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aempty[buf]);
      const float* bs_s = colp + cb;
      const float* bt_s = colp + ncols + cb;
      float out[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float s = fmaf(sv[e], c2s, bs_s[e]), t = fmaf(tv[e], c2t, bt_s[e]), zv = z16[e];
        const bool keep = (mb >> e) & 1u;
        const float u = ex2_fast(-1.4426950408889634f * fabsf(s));  // exp(-|s|) in (0, 1]
        const float rcp = rcp_fast(1.f + u);                        // sigmoid(|s|) in [1/2, 1)
        const float gate = s >= 0.f ? rcp : u * rcp;                // sigmoid(s)
        const float lg = fmaf(lg2_fast(rcp), 0.6931471805599453f, fminf(s, 0.f));  // log(sigmoid(s))
        out[e] = fmaf(1.f - gate, t, keep ? zv : zv * gate);        // (1 - m) z gate + (1 - gate) t + m z
        ld += keep ? 0.f : lg;
        zm = fmaxf(zm, fabsf(out[e]));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(zt + (((q0 + q) ^ sw) << 4)) = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
      fence_proxy_async();
      named_sync(1 + par, 256);  // both column halves of the tile are in place
      if (half == 0 && g == 0) {  // one thread stores the tile and hands the slot back to the loader
        if (lane == 0) {
          tma_store_2d(&tm_out, c * TK, (int)row0, raw + slot * RAW_TILE);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive_n(&raw_free[slot], 4);
        }
        __syncwarp();
      }
    }
    ldp[(half + 2 * par) * BM + r] = ld;
    if (!row_ok) zm = 0.f;  // rows beyond B hold zeros on input but not on output
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) zm = fmaxf(zm, __shfl_xor_sync(0xffffffffu, zm, o));
  if (lane == 0) zmx[warp] = zm;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < BM) {
    const int64_t row = row0 + threadIdx.x;
    if (row < a.B && a.ld_out)
      a.ld_out[row] = (a.ld_in ? a.ld_in[row] : 0.f) + ((ldp[threadIdx.x] + ldp[BM + threadIdx.x]) + (ldp[2 * BM + threadIdx.x] + ldp[3 * BM + threadIdx.x]));
  }
  if (threadIdx.x == 0 && a.tmax_out) {
    float m = 0.f;
    for (int w = 0; w < WARPS; ++w) m = fmaxf(m, zmx[w]);
    a.tmax_out[blockIdx.x] = m;
  }
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---- weights.  One power-of-two scale per matrix (max |W| * 2^e in [2^13, 2^14)): elements far below the
// maximum lose relative precision in their lo plane (fp16 subnormals) but stay within 2^-38 of the maximum in
// absolute terms.  Block 0: W1, block 1: Ws, block 2: Wt.
__global__ void __launch_bounds__(1024) mat_scales(int D, int H, const float* __restrict__ w1, const float* __restrict__ ws,
                                                   const float* __restrict__ wt, int64_t ld2, int* __restrict__ esc,
                                                   float* __restrict__ sinv) {
  __shared__ float red[32];
  float m = 0.f;
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < D * H; i += 1024) m = fmaxf(m, fabsf(w1[i]));
  } else {
    const float* w = blockIdx.x == 1 ? ws : wt;
    for (int i = threadIdx.x; i < D * H; i += 1024) {
      const int h = i / D, c = i - h * D;
      m = fmaxf(m, fabsf(w[(int64_t)h * ld2 + c]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w) m = fmaxf(m, red[w]);
    const int e = scale_exp(m);
    esc[blockIdx.x] = e;
    sinv[blockIdx.x] = ldexpf(1.f, -e);
  }
}
// one thread = one 16-byte chunk (8 consecutive k of one column), hi and lo
__global__ void pack_weights(int D, int H, int kblocks, const float* __restrict__ w1, const float* __restrict__ ws,
                             const float* __restrict__ wt, int64_t ld2, const int* __restrict__ esc,
                             uint8_t* __restrict__ W1p, uint8_t* __restrict__ W2p) {
  const int64_t n1 = (int64_t)kblocks * (TK / 8) * HP, n2 = (int64_t)kblocks * (HP / 8) * 64;
  const int e1 = esc[0], es = esc[1], et = esc[2];
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n1 + n2; idx += (int64_t)gridDim.x * blockDim.x) {
    float w[8];
    uint8_t* dst;
    int vstride;
    if (idx < n1) {
      const int n = (int)(idx % HP);
      int64_t t = idx / HP;
      const int kc = (int)(t % (TK / 8)), kb = (int)(t / (TK / 8));
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = kb * TK + kc * 8 + e;
        w[e] = (k < D && n < H) ? ldexpf(w1[(int64_t)k * H + n], e1) : 0.f;
      }
      dst = W1p + (int64_t)kb * B1_STAGE + ((int64_t)kc * HP + n) * 16;
      vstride = B1_VAR;
    } else {
      const int64_t i2 = idx - n1;
      const int n = (int)(i2 % 64);
      int64_t t = i2 / 64;
      const int kc = (int)(t % (HP / 8)), c = (int)(t / (HP / 8));
      const int col = c * TK + (n & 31);
      const float* wsrc = n < 32 ? ws : wt;
      const int e2 = n < 32 ? es : et;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int h = kc * 8 + e;
        w[e] = (h < H && col < D) ? ldexpf(wsrc[(int64_t)h * ld2 + col], e2) : 0.f;
      }
      dst = W2p + (int64_t)c * B2_STAGE + ((int64_t)kc * 64 + n) * 16;
      vstride = B2_VAR;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_f16x2(w[2 * e], w[2 * e + 1], hi[e], lo[e]);
    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst + vstride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- z0 = q0_mean + exp(q0_log_var / 2) * eps for all rows (mnf_dense.py:93-96) plus the max |z0| of every
// 128-row tile (atomicMax on the non-negative float's bit pattern; tmax zeroed by the caller).  A CTA owns ZR
// rows of one tile; a thread draws four consecutive elements per Philox call (same counters as every other
// consumer of this draw: row, inner = d >> 2).
constexpr int ZR = 8;
template <bool INJ>  // INJ: eps read from memory (tests), compiled apart: the Philox path is branch-free
__global__ void __launch_bounds__(256) sample_z0_tmax(int64_t B, int D, const float* __restrict__ q0_mean,
                                                      const float* __restrict__ q0_log_var, NoiseDev eps,
                                                      float* __restrict__ zout, float* __restrict__ tmax) {
  __shared__ float red[8];
  extern __shared__ __align__(16) float zsm[];  // [2][D4 * 4]: q0_mean | exp(q0_log_var / 2), once per CTA (not per element)
  const int D4 = (D + 3) >> 2;
  float* mean_s = zsm;
  float* sd_s = zsm + 4 * D4;
  for (int d = threadIdx.x; d < 4 * D4; d += 256) {
    mean_s[d] = d < D ? __ldg(q0_mean + d) : 0.f;
    sd_s[d] = d < D ? expf(0.5f * __ldg(q0_log_var + d)) : 0.f;
  }
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * ZR;
  const PhiloxKeys K = philox_keys(eps);
  float m = 0.f;
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(zout) & 15) == 0);
  // (row, quad) of this thread's items, advanced incrementally (no division per item)
  const int step_r = 256 / D4, step_g = 256 - step_r * D4;
  int rr = (int)threadIdx.x / D4, g4 = (int)threadIdx.x - rr * D4;
  for (; rr < ZR; rr += step_r, g4 += step_g) {
    if (g4 >= D4) { g4 -= D4; ++rr; if (rr >= ZR) break; }
    const int d0 = g4 * 4;
    const int64_t r = r0 + rr;
    if (r >= B) break;
    float nrm[4];
    if (INJ) {
#pragma unroll
      for (int e = 0; e < 4; ++e) nrm[e] = d0 + e < D ? eps.injected[r * D + d0 + e] : 0.f;
    } else {
      philox_normal4_keys(eps, K, eps.row_offset + (uint64_t)r, 0u, (uint32_t)g4, nrm);
    }
    const float4 mu = *reinterpret_cast<const float4*>(mean_s + d0), sd = *reinterpret_cast<const float4*>(sd_s + d0);
    const float v[4] = {fmaf(sd.x, nrm[0], mu.x), fmaf(sd.y, nrm[1], mu.y), fmaf(sd.z, nrm[2], mu.z), fmaf(sd.w, nrm[3], mu.w)};
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v[0]), fabsf(v[1]))), fmaxf(fabsf(v[2]), fabsf(v[3])));  // (padding elements are 0)
    if (vec) *reinterpret_cast<float4*>(zout + r * D + d0) = make_float4(v[0], v[1], v[2], v[3]);
    else
      for (int e = 0; e < 4; ++e)
        if (d0 + e < D) zout[r * D + d0 + e] = v[e];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    if (m == m) atomicMax(reinterpret_cast<int*>(tmax + r0 / BM), __float_as_int(m));
  }
}
// max |z| per 128-row tile of an existing z (chains that start from a caller's z0)
__global__ void __launch_bounds__(256) tile_absmax(int64_t B, int D, const float* __restrict__ z, float* __restrict__ tmax) {
  __shared__ float red[8];
  const int64_t r0 = (int64_t)blockIdx.x * ZR;
  const int64_t n = (B - r0 < ZR ? B - r0 : ZR) * (int64_t)D;
  const float* p = z + r0 * D;
  float m = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) m = fmaxf(m, fabsf(__ldg(p + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    if (m == m) atomicMax(reinterpret_cast<int*>(tmax + r0 / BM), __float_as_int(m));
  }
}

}  // namespace f2

// ------------------------------------------------------------------------------------------ host side
bool mnf_flow_f16_ok(const mnf_ctx* ctx, int64_t B, int64_t D, const mnf_flow_step& st) {
  static const bool off = getenv("MNF_NO_FLOW_F16") != nullptr;  // A/B switch: the 3xTF32 kernel of flows_umma.cu
  return !off && ctx->math_mode != MNF_MATH_FP32 && st.kind == MNF_FLOW_RNVP && !mnf_flow_step_is_generic(st) && B >= 32 &&
         D >= 4 && D <= 800 && D % 4 == 0 && st.hidden >= 1 && st.hidden <= f2::HP && !st.m1 && !st.m2;
}
size_t mnf_flow_f16_pack_bytes(int D) {
  const size_t kb = (size_t)(D + f2::TK - 1) / f2::TK;
  return kb * (f2::B1_STAGE + f2::B2_STAGE) + 256;
}
size_t mnf_flow_f16_tiles(int64_t B) { return (size_t)cdiv(B, f2::BM); }

// tmax[tiles] = 0, then z0 for all rows + the tile maxima
int mnf_flow_f16_sample_z0(mnf_ctx* ctx, int64_t B, int D, const float* q0_mean, const float* q0_log_var, const NoiseDev& eps,
                           float* zout, float* tmax) {
  MNF_CUDA(cudaMemsetAsync(tmax, 0, sizeof(float) * mnf_flow_f16_tiles(B), ctx->stream));
  const size_t zsm = sizeof(float) * 8 * ((D + 3) / 4);
  if (eps.injected) f2::sample_z0_tmax<true><<<(unsigned)cdiv(B, f2::ZR), 256, zsm, ctx->stream>>>(B, D, q0_mean, q0_log_var, eps, zout, tmax);
  else f2::sample_z0_tmax<false><<<(unsigned)cdiv(B, f2::ZR), 256, zsm, ctx->stream>>>(B, D, q0_mean, q0_log_var, eps, zout, tmax);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}
int mnf_flow_f16_absmax(mnf_ctx* ctx, int64_t B, int D, const float* z, float* tmax) {
  MNF_CUDA(cudaMemsetAsync(tmax, 0, sizeof(float) * mnf_flow_f16_tiles(B), ctx->stream));
  f2::tile_absmax<<<(unsigned)cdiv(B, f2::ZR), 256, 0, ctx->stream>>>(B, D, z, tmax);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

// TMA descriptor of a row-major [B, D] fp32 matrix, box = [128 rows x 32 columns], SWIZZLE_128B.  The driver entry
// point comes from the runtime (no link against libcuda).
typedef CUresult (*mnf_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_tile_map(CUtensorMap* tm, const float* base, int64_t B, int D) {
  static mnf_tmap_encode_fn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    MNF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    MNF_CHECK_ARG(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
    enc = (mnf_tmap_encode_fn)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)B};
  const cuuint64_t strides[1] = {(cuuint64_t)D * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)f2::TK, (cuuint32_t)f2::BM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MNF_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for a [%lld, %d] matrix", (int)r, (long long)B, D);
  return MNF_OK;
}

// One RNVP step; pack_ws: mnf_flow_f16_pack_bytes(D) bytes of workspace (ignored when the weights are frozen).
int mnf_flow_f16_step(mnf_ctx* ctx, int64_t B, int D, const float* zin, float* zout, const float* ld_in, float* ld_out,
                      float* hid_save, const mnf_flow_step& st, const NoiseDev& bern, const float* tmax_in, float* tmax_out,
                      uint8_t* pack_ws) {
  using namespace f2;
  MNF_CHECK_ARG(((reinterpret_cast<uintptr_t>(zin) | reinterpret_cast<uintptr_t>(zout)) & 15) == 0, "flow_f16: z buffers must be 16-byte aligned");
  Args a = {};
  a.B = B; a.D = D; a.H = st.hidden; a.kblocks = (D + TK - 1) / TK;
  a.zin = zin; a.zout = zout; a.ld_in = ld_in; a.ld_out = ld_out; a.hid_save = hid_save;
  a.b1 = st.b1; a.bs = st.bs; a.bt = st.bt; a.bern = bern; a.tmax_in = tmax_in; a.tmax_out = tmax_out;
  bool fresh = true;
  if (ctx->frozen) {
    pack_ws = (uint8_t*)mnf_packed(ctx, mnf_mix(0xF16F2ull, ((uint64_t)D << 32) | (uint32_t)st.hidden), (uint64_t)(uintptr_t)st.w1,
                                   mnf_mix((uint64_t)(uintptr_t)st.ws, (uint64_t)(uintptr_t)st.wt), (uint64_t)st.ld2,
                                   mnf_flow_f16_pack_bytes(D), &fresh);
    if (!pack_ws) return MNF_ERR_CUDA;
  }
  uint8_t* W1p = pack_ws;
  uint8_t* W2p = W1p + (size_t)a.kblocks * B1_STAGE;
  float* sinv = (float*)(W2p + (size_t)a.kblocks * B2_STAGE);
  int* esc = (int*)(sinv + 4);
  if (fresh) {
    mat_scales<<<3, 1024, 0, ctx->stream>>>(D, st.hidden, st.w1, st.ws, st.wt, st.ld2, esc, sinv);
    MNF_LAUNCHED(ctx);
    const int64_t tot = (int64_t)a.kblocks * ((TK / 8) * HP + (HP / 8) * 64);
    pack_weights<<<(unsigned)cdiv(tot, 128), 128, 0, ctx->stream>>>(D, st.hidden, a.kblocks, st.w1, st.ws, st.wt, st.ld2, esc, W1p, W2p);
    MNF_LAUNCHED(ctx);
  }
  a.W1p = W1p; a.W2p = W2p; a.sinv = sinv;
  CUtensorMap tm_in, tm_out;
  if (int rc = make_tile_map(&tm_in, zin, B, D)) return rc;
  if (int rc = make_tile_map(&tm_out, zout, B, D)) return rc;
  const size_t smem = (size_t)REGION_X + NR * RAW_TILE + sizeof(uint32_t) * (size_t)a.kblocks * BM +
                      sizeof(float) * (2 * (size_t)a.kblocks * TK + HP + 4 * BM + 32) + 8 * 32 + 16;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MNF_CUDA(cudaFuncSetAttribute(rnvp_f16_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const bool probed = mnf_probe_begin(ctx, MNF_PROBE_FLOW_STEP);
  mnf_note_path("flow_f16");
  rnvp_f16_step<<<(unsigned)cdiv(B, BM), THREADS, smem, ctx->stream>>>(a, tm_in, tm_out);
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}
