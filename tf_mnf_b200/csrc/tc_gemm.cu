// Tensor-core path of the paired mean/variance contraction (MNFDense.call mnf_dense.py:161-166,
// MNFConv2D.call mnf_conv.py:190-191) for sm_100a: tcgen05.mma (kind::tf32) with TMEM
// accumulators, warp-specialised.
//
// fp32 parity on tensor cores: every operand is split x = hi + lo (both rounded to nearest
// TF32 values) and each product is evaluated as hi*hi + lo*hi + hi*lo (3xTF32) with fp32
// accumulation in TMEM: per-term relative error ~2^-20, inside the 1e-5 bar.
//
// One CTA per SM, persistent over 128-row tiles:
//   warps 0-3  epilogue : tcgen05.ld the mean/var accumulators, bias, conv z-scale, Philox eps,
//                         stage y in shared memory, coalesced store
//   warp  4    MMA      : one elected thread issues 6 tcgen05.mma per 8-wide k-step
//                         (3 split terms x {mean, var}); tcgen05.commit frees smem stages
//   warps 5-12 producer : gather the x tile ONCE (im2col on the fly for conv), form
//                         {x(*z), x^2} x {hi, lo} in registers and write them K-major into the
//                         no-swizzle UMMA layout; the pre-split weight block arrives by one
//                         cp.async.bulk (TMA, 1-D) per stage.
#include <cstdlib>

#include "common.cuh"

namespace tc {

constexpr int BM = 128;
constexpr int BK = 16;       // floats per pipeline stage
constexpr int KC = BK / 4;   // 16-byte k-chunks per stage
// BN = 128 (wide dense layers): 64 KB per stage and no output staging tile (rows are stored straight
// from registers), so that three stages fit: with two, the weight block of the next stage is still
// on its way from L2 when the MMAs of the current one retire (62 % of the time waiting, measured)
constexpr int stages_for(int bn) { return 3 + 0 * bn; }
constexpr bool direct_out(int bn, bool conv) { return bn > 64 && !conv; }
constexpr int EPI_WARPS = 4, PROD_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;  // 416
constexpr int A_VARIANT_BYTES = KC * BM * 16;               // 8 KB
constexpr int A_STAGE_BYTES = 4 * A_VARIANT_BYTES;          // 32 KB

struct Args {
  int64_t M;
  int K, N;
  int kblocks;       // ceil(K / BK)
  int ntiles;        // ceil(N / BN)
  const float* x;
  const float* z;    // dense [M,K] / conv [B,N] / nullptr
  const float* Bp;   // packed split weights: [ntile][kblock][4 variants][KC][BN][4]
  const float* bm;   // [N]
  const float* vb;   // [N]
  float* y;
  float* mean_out;
  float* sd_out;
  NoiseDev eps;
  int H, W, C, kh, kw, Ho, Wo, stride, pad_t, pad_l;
  int vec;           // 16-byte gathers legal (conv: C % 4 == 0; dense: K % 4 == 0)
  int act;           // dense: 1 = ReLU fused into the epilogue
  int pool;          // conv: ReLU + 2x2 max-pool fused into the copy-out (ntiles == 1, BM % (2 Wo) == 0)
};

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // Bounded: a protocol error (a barrier nobody will ever complete) must end as a launch failure the host
  // sees, not as a kernel that spins forever and takes the GPU with it.  try_wait blocks in hardware for a
  // while before it returns false, so 2^24 failed polls are many seconds -- never reached by a live pipeline.
  uint32_t done = 0;
  for (uint32_t polls = 0; !done; ++polls) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && polls > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//  [0,14) start >> 4 | [16,30) leading byte offset >> 4 (between the two 16-byte K chunks of
//  one MMA) | [32,46) stride byte offset >> 4 (between 8-row groups) | [46,48) version = 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Warp-converged issue: all lanes execute, one elected lane performs the operation.  Issued from
// `if (lane == 0)` the compiler treats every operand as a per-thread value and each tcgen05.mma
// costs an R2UR chain inside an ELECT / BRA.U.ANY loop (~100 cycles per instruction measured);
// from converged code with warp-uniform operands the descriptors live in uniform registers.
__device__ __forceinline__ void umma_tf32_e(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u));
}
__device__ __forceinline__ void umma_commit_e(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// x = hi + lo with BOTH parts rounded to nearest TF32 (cvt.rna): truncation would bias every
// term the same way and the error of a K-term sum would grow like K instead of sqrt(K)
__device__ __forceinline__ float tf32_rna(float a) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = tf32_rna(a);
  lo = tf32_rna(a - hi);
}

// ---------------------------------------------------------------------------------- kernel
template <int BN, bool CONV>
__global__ void __launch_bounds__(THREADS, 1) paired_umma_fwd(Args a) {
  constexpr int STAGES = stages_for(BN);
  constexpr bool DIRECT = direct_out(BN, CONV);
  constexpr int B_VARIANT_BYTES = KC * BN * 16;
  constexpr int B_STAGE_BYTES = 4 * B_VARIANT_BYTES;
  constexpr int TMEM_COLS = 4 * BN;  // 2 buffers x {mean, var}
  constexpr int LDY = BN + 1;
  static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;                                  // [STAGES][4][KC][BM][4 floats]
  uint8_t* Bs = As + STAGES * A_STAGE_BYTES;           // [STAGES][4][KC][BN][4 floats]
  float* ys = (float*)(Bs + STAGES * B_STAGE_BYTES);   // [BM][LDY]
  uint64_t* bars = (uint64_t*)(ys + (DIRECT ? 0 : BM * LDY) + 3);
  bars = (uint64_t*)(((uintptr_t)bars + 7) & ~(uintptr_t)7);
  uint64_t* full = bars;               // [STAGES]
  uint64_t* empty = bars + STAGES;     // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES; // [2]
  uint64_t* tempty = tfull + 2;        // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  float* bias_s = (float*)(tmem_slot + 2);   // [2][BN]: mean_b, var_b of the current n-tile (ntiles == 1 only)
  int* ktab = (int*)(bias_s + 2 * BN);       // conv gather table, one entry per k-chunk (vec) or per k

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int64_t mtiles = (a.M + BM - 1) / BM;
  const int64_t ntot = mtiles * a.ntiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      // dense (vector path): all 8 producer warps fill every stage; otherwise 4 warps of one
      // producer group; plus the B bulk-copy arrive
      mbar_init(&full[s], ((!CONV && a.vec) ? PROD_WARPS : PROD_WARPS / 2) + 1);
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
  // im2col table: offset of k (or of the 4-channel chunk starting at k) relative to the row's
  // patch origin, tap (i,j) in the high bits; -1 = beyond K (zero fill)
  if (CONV) {
    const int gran = a.vec ? 4 : 1;
    const int nent = a.kblocks * BK / gran;
    const int kwC = a.kw * a.C;
    for (int q = threadIdx.x; q < nent; q += THREADS) {
      const int k = q * gran;
      int v = -1;
      if (k < a.K) {
        const int i = k / kwC, rem = k - i * kwC, j = rem / a.C, cch = rem - j * a.C;
        v = (i * a.W + j) * a.C + cch;
        ktab[nent + q] = (i << 16) | j;
      }
      ktab[q] = v;
    }
  }
  if (a.ntiles == 1)
    for (int n = threadIdx.x; n < BN; n += THREADS) {
      bias_s[n] = n < a.N ? a.bm[n] : 0.f;
      bias_s[BN + n] = n < a.N ? a.vb[n] : 1.f;
    }
  __syncthreads();

  if (warp < EPI_WARPS) {
    // ================================================================== epilogue warps
    const int r = warp * 32 + lane;  // tile row == TMEM lane
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntot; t += gridDim.x, ++it) {
      const int64_t mt = t / a.ntiles;
      const int nt = (int)(t - mt * a.ntiles);
      const int64_t m0 = mt * BM, m = m0 + r;
      const int n0 = nt * BN;
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&tfull[buf], ph);
      tc_fence_after();
      const bool row_ok = m < a.M;
      const int ncols = (a.N - n0 < BN) ? a.N - n0 : BN;
      const int ldy = (a.ntiles == 1) ? a.N : LDY;
      const int64_t HoWo = CONV ? (int64_t)a.Ho * a.Wo : 1;
      const int64_t brow = CONV ? m / HoWo : m;
      const uint32_t outer = CONV ? (uint32_t)(m - brow * HoWo) : 0u;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * 2 * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        if (n0 + c0 >= a.N) break;  // warp-uniform
        float mv[16], vv[16];
        tmem_ld16(trow + c0, mv);
        tmem_ld16(trow + BN + c0, vv);
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int nb = n0 + c0 + 4 * q;
            if (nb >= a.N) break;
            float e[4];
            if (a.eps.injected) {
#pragma unroll
              for (int j = 0; j < 4; ++j) e[j] = nb + j < a.N ? a.eps.injected[m * a.N + nb + j] : 0.f;
            } else {
              philox_normal4(a.eps, a.eps.row_offset + (uint64_t)brow, outer, (uint32_t)(nb >> 2), e);
            }
            float o4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = nb + j;
              if (n < a.N) {
                const float bmn = (a.ntiles == 1) ? bias_s[n] : __ldg(a.bm + n);
                const float vbn = (a.ntiles == 1) ? bias_s[BN + n] : __ldg(a.vb + n);
                float mean = mv[4 * q + j] + bmn;
                float sd = sqrtf(vv[4 * q + j] + vbn);
                if (a.sd_out) a.sd_out[m * a.N + n] = sd;
                if (CONV) {
                  if (a.mean_out) a.mean_out[m * a.N + n] = mean;
                  if (a.z) mean *= __ldg(a.z + brow * a.N + n);
                }
                const float out = fmaf(sd, e[j], mean);
                const float outa = (!CONV && a.act == 1) ? fmaxf(out, 0.f) : out;
                if (DIRECT) o4[j] = outa;
                else ys[r * ldy + c0 + 4 * q + j] = outa;
              }
            }
            if (DIRECT) {  // 16 contiguous bytes of this thread's output row
              float* dstp = a.y + m * a.N + nb;
              if (nb + 3 < a.N && !(a.N & 3)) *reinterpret_cast<float4*>(dstp) = make_float4(o4[0], o4[1], o4[2], o4[3]);
              else
                for (int j = 0; j < 4; ++j)
                  if (nb + j < a.N) dstp[j] = o4[j];
            }
          }
        }
      }
      // accumulators consumed: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
      if (DIRECT) continue;  // rows already stored from registers
      // coalesced copy-out of the staged tile (rows are contiguous in y when ntiles == 1)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int rows = (int)((a.M - m0 < BM) ? a.M - m0 : BM);
      if (a.pool) {
        // rows of the tile are consecutive output pixels and the tile starts on an even output
        // row, so its BM/4 pooled pixels are one contiguous block of the pooled tensor
        const int Wp = a.Wo >> 1, np = rows >> 2;
        float* dst = a.y + (m0 >> 2) * a.N;
        for (int e = threadIdx.x; e < np * a.N; e += EPI_WARPS * 32) {
          const int pp = e / a.N, c = e - pp * a.N;
          const int hp = pp / Wp, wp = pp - hp * Wp;
          const float* s0 = ys + ((2 * hp) * a.Wo + 2 * wp) * ldy + c;
          float v = fmaxf(fmaxf(s0[0], s0[ldy]), fmaxf(s0[a.Wo * ldy], s0[(a.Wo + 1) * ldy]));
          dst[e] = fmaxf(v, 0.f);
        }
      } else if (a.ntiles == 1) {  // the tile is one contiguous block of y and is staged densely
        float* dst = a.y + m0 * a.N;
        for (int e = threadIdx.x; e < rows * a.N; e += EPI_WARPS * 32) dst[e] = ys[e];
      } else {
        for (int e = threadIdx.x; e < rows * ncols; e += EPI_WARPS * 32) {
          int rr = e / ncols, cc = e - rr * ncols;
          a.y[(m0 + rr) * a.N + n0 + cc] = ys[rr * LDY + cc];
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  } else if (warp == EPI_WARPS) {
    // ======================================================================= MMA issuer
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, K-major both
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint64_t adesc0 = make_desc(smem_u32(As), BM * 16, 128), bdesc0 = make_desc(smem_u32(Bs), BN * 16, 128);
    int it = 0, stage = 0;
    uint32_t sph = 0;
    for (int64_t t = blockIdx.x; t < ntot; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&tempty[buf], ph ^ 1);
      tc_fence_after();
      const uint32_t d_mean = tmem_base + (uint32_t)(buf * 2 * BN), d_var = d_mean + BN;
      for (int kb = 0; kb < a.kblocks; ++kb) {
        mbar_wait(&full[stage], sph);
        tc_fence_after();
        {
          // descriptors differ only in their 14-bit start-address field: add offsets to a base
          const uint64_t ad = adesc0 + (uint64_t)((stage * A_STAGE_BYTES) >> 4);
          const uint64_t bd = bdesc0 + (uint64_t)((stage * B_STAGE_BYTES) >> 4);
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t ao = ad + (uint64_t)((2 * j * BM * 16) >> 4), bo = bd + (uint64_t)((2 * j * BN * 16) >> 4);
            const uint64_t xh = ao, xl = ao + (A_VARIANT_BYTES >> 4), qh = ao + (2 * A_VARIANT_BYTES >> 4),
                           ql = ao + (3 * A_VARIANT_BYTES >> 4);
            const uint64_t wh = bo, wl = bo + (B_VARIANT_BYTES >> 4), vh = bo + (2 * B_VARIANT_BYTES >> 4),
                           vl = bo + (3 * B_VARIANT_BYTES >> 4);
            const uint32_t first = (kb == 0 && j == 0) ? 0u : 1u;
            umma_tf32_e(d_mean, xh, wh, idesc, first);
            umma_tf32_e(d_mean, xl, wh, idesc, 1u);
            umma_tf32_e(d_mean, xh, wl, idesc, 1u);
            umma_tf32_e(d_var, qh, vh, idesc, first);
            umma_tf32_e(d_var, ql, vh, idesc, 1u);
            umma_tf32_e(d_var, qh, vl, idesc, 1u);
          }
          umma_commit_e(&empty[stage]);                          // smem stage free once these MMAs retire
          if (kb == a.kblocks - 1) umma_commit_e(&tfull[buf]);   // accumulators ready
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; sph ^= 1; }
      }
    }
  } else {
    // ======================================================================== producers
    // Two groups of 4 warps; group g stages the k-blocks with (global index) % 2 == g, so two
    // pipeline stages are always being filled concurrently.  A thread owns one tile row and
    // the 16 k of a block: its four 16-byte gathers are issued together and the NEXT block's
    // gathers are in flight while the current block is split and written to shared memory.
    const int pt = threadIdx.x - (EPI_WARPS + 1) * 32;  // 0..255
    if (!CONV && a.vec) {
      // ---- dense rows, K % 4 == 0: the global-load latency (L2, ~1 us under load) is the
      // enemy.  Four lanes read the 64 contiguous bytes of one row's k-block (coalesced), every
      // thread owns two rows x one 16-byte k-chunk, and the loads of the next PF k-blocks are
      // kept in flight in a register ring while the current block is split and stored.
      constexpr int PF = 4;
      const int kc = pt & 3, rs = pt >> 2;  // rows rs and rs + 64
      const int my_tiles = (int)((ntot > blockIdx.x) ? (ntot - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
      const int64_t nblk = (int64_t)my_tiles * a.kblocks;
      float4 xb[PF][2], zb[PF][2];
      // load cursor
      int lti = 0, lkb = 0;
      const float *lx0 = nullptr, *lx1 = nullptr, *lz0 = nullptr, *lz1 = nullptr;
      auto load_tile = [&](int ti) {
        const int64_t t = blockIdx.x + (int64_t)ti * gridDim.x;
        const int64_t mt = t / a.ntiles;
        int64_t m0 = mt * BM + rs, m1 = m0 + 64;
        if (m0 >= a.M) m0 = a.M - 1;  // rows beyond M read a valid row; never stored
        if (m1 >= a.M) m1 = a.M - 1;
        lx0 = a.x + m0 * a.K + kc * 4;
        lx1 = a.x + m1 * a.K + kc * 4;
        if (a.z) { lz0 = a.z + m0 * a.K + kc * 4; lz1 = a.z + m1 * a.K + kc * 4; }
      };
      auto issue = [&](float4 (&xv)[2], float4 (&zv)[2]) {
        if (lkb == 0) load_tile(lti);
        const int k = lkb * BK + kc * 4;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
        xv[0] = xv[1] = zero;
        zv[0] = zv[1] = one;
        if (k < a.K) {
          xv[0] = __ldg(reinterpret_cast<const float4*>(lx0 + lkb * BK));
          xv[1] = __ldg(reinterpret_cast<const float4*>(lx1 + lkb * BK));
          if (a.z) {
            zv[0] = __ldg(reinterpret_cast<const float4*>(lz0 + lkb * BK));
            zv[1] = __ldg(reinterpret_cast<const float4*>(lz1 + lkb * BK));
          }
        }
        if (++lkb == a.kblocks) { lkb = 0; ++lti; }
      };
#pragma unroll
      for (int d = 0; d < PF; ++d)
        if (d < nblk) issue(xb[d], zb[d]);
      int sti = 0, skb = 0, stage = 0;
      uint32_t sph = 0;
      for (int64_t b0 = 0; b0 < nblk; b0 += PF) {
#pragma unroll
        for (int d = 0; d < PF; ++d) {
          const int64_t b = b0 + d;
          if (b < nblk) {
            mbar_wait(&empty[stage], sph ^ 1);
            if (pt == 0) {
              const int64_t t = blockIdx.x + (int64_t)sti * gridDim.x;
              const int nt = (int)(t % a.ntiles);
              const float* bsrc = a.Bp + ((size_t)nt * a.kblocks + skb) * (size_t)(B_STAGE_BYTES / 4);
              mbar_arrive_expect_tx(&full[stage], B_STAGE_BYTES);
              bulk_g2s(Bs + stage * B_STAGE_BYTES, bsrc, B_STAGE_BYTES, &full[stage]);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float xs[4] = {xb[d][h].x, xb[d][h].y, xb[d][h].z, xb[d][h].w};
              const float zs[4] = {zb[d][h].x, zb[d][h].y, zb[d][h].z, zb[d][h].w};
              float h1[4], l1[4], h2[4], l2[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                split_tf32(xs[e] * zs[e], h1[e], l1[e]);
                split_tf32(xs[e] * xs[e], h2[e], l2[e]);
              }
              uint8_t* dst = As + stage * A_STAGE_BYTES + kc * (BM * 16) + (rs + 64 * h) * 16;
              *reinterpret_cast<float4*>(dst + 0 * A_VARIANT_BYTES) = make_float4(h1[0], h1[1], h1[2], h1[3]);
              *reinterpret_cast<float4*>(dst + 1 * A_VARIANT_BYTES) = make_float4(l1[0], l1[1], l1[2], l1[3]);
              *reinterpret_cast<float4*>(dst + 2 * A_VARIANT_BYTES) = make_float4(h2[0], h2[1], h2[2], h2[3]);
              *reinterpret_cast<float4*>(dst + 3 * A_VARIANT_BYTES) = make_float4(l2[0], l2[1], l2[2], l2[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
            if (b + PF < nblk) issue(xb[d], zb[d]);  // refill this ring slot
            if (++skb == a.kblocks) { skb = 0; ++sti; }
            if (++stage == STAGES) { stage = 0; sph ^= 1; }
          }
        }
      }
    } else {
    const int r = pt & (BM - 1);
    const int grp = pt >> 7;
    struct RowCtx {
      int tile_i;       // which of this CTA's tiles the cached row info belongs to
      int64_t m;
      const float* xrow;  // conv: x + rowoff ; dense: x + m*K
      const float* zrow;  // dense: z + m*K
      int hi0, wi0, nt;
      bool ok;
    } rc;
    rc.tile_i = -1;
    auto set_tile = [&](int ti) {
      if (rc.tile_i == ti) return;
      rc.tile_i = ti;
      const int64_t t = blockIdx.x + (int64_t)ti * gridDim.x;
      const int64_t mt = t / a.ntiles;
      rc.nt = (int)(t - mt * a.ntiles);
      rc.m = mt * BM + r;
      rc.ok = rc.m < a.M;
      if (!rc.ok) rc.m = a.M - 1;  // gather from a valid row; the epilogue never stores it
      rc.hi0 = rc.wi0 = 0;
      rc.xrow = a.x;
      rc.zrow = a.z;
      {
        if (CONV) {
          const int64_t HoWo = (int64_t)a.Ho * a.Wo;
          const int64_t b = rc.m / HoWo;
          const int pix = (int)(rc.m - b * HoWo);
          const int ho = pix / a.Wo, wo = pix - ho * a.Wo;
          rc.hi0 = ho * a.stride - a.pad_t;
          rc.wi0 = wo * a.stride - a.pad_l;
          rc.xrow = a.x + ((b * a.H + rc.hi0) * a.W + rc.wi0) * a.C;
        } else {
          rc.xrow = a.x + rc.m * a.K;
          if (a.z) rc.zrow = a.z + rc.m * a.K;
        }
      }
    };
    const bool padded = CONV && (a.pad_t != 0 || a.pad_l != 0 || (a.Ho - 1) * a.stride + a.kh > a.H ||
                                 (a.Wo - 1) * a.stride + a.kw > a.W);
    const int nent = CONV ? a.kblocks * BK / (a.vec ? 4 : 1) : 0;
    // gather the 16 x values (and z for dense) of k-block kb of this CTA's tile ti into registers.
    // Rows beyond M read a valid row (their results are never stored), so the fast paths carry no
    // row predicate; conv offsets come from the shared-memory table.
    auto gather = [&](int ti, int kb, float xv[16], float zv[16]) {
      set_tile(ti);
      const int k0 = kb * BK;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        const int k = k0 + kc * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) { xv[kc * 4 + e] = 0.f; zv[kc * 4 + e] = 1.f; }
        if (CONV) {
          if (a.vec) {
            const int off = ktab[(k0 >> 2) + kc];
            bool ok = off >= 0;
            if (padded && ok) {
              const int ij = ktab[nent + (k0 >> 2) + kc];
              const int hi = rc.hi0 + (ij >> 16), wi = rc.wi0 + (ij & 0xFFFF);
              ok = hi >= 0 && hi < a.H && wi >= 0 && wi < a.W;
            }
            if (ok) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(rc.xrow + off));
              xv[kc * 4 + 0] = v.x; xv[kc * 4 + 1] = v.y; xv[kc * 4 + 2] = v.z; xv[kc * 4 + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int off = ktab[k + e];
              bool ok = off >= 0;
              if (padded && ok) {
                const int ij = ktab[nent + k + e];
                const int hi = rc.hi0 + (ij >> 16), wi = rc.wi0 + (ij & 0xFFFF);
                ok = hi >= 0 && hi < a.H && wi >= 0 && wi < a.W;
              }
              if (ok) xv[kc * 4 + e] = __ldg(rc.xrow + off);
            }
          }
        } else if (k < a.K) {
          if (a.vec) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(rc.xrow + k));
            xv[kc * 4 + 0] = v.x; xv[kc * 4 + 1] = v.y; xv[kc * 4 + 2] = v.z; xv[kc * 4 + 3] = v.w;
            if (a.z) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(rc.zrow + k));
              zv[kc * 4 + 0] = w.x; zv[kc * 4 + 1] = w.y; zv[kc * 4 + 2] = w.z; zv[kc * 4 + 3] = w.w;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (k + e < a.K) {
                xv[kc * 4 + e] = __ldg(rc.xrow + k + e);
                if (a.z) zv[kc * 4 + e] = __ldg(rc.zrow + k + e);
              }
          }
        }
      }
    };

    // group g walks the k-blocks c = g, g+2, ... of this CTA's (tile, k-block) sequence
    const int my_tiles = (int)((ntot > blockIdx.x) ? (ntot - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
    int ti = 0, kb = grp, stage = grp;
    uint32_t sph = 0;
    while (kb >= a.kblocks) { kb -= a.kblocks; ++ti; }
    int nti = ti, nkb = kb;
    auto advance2 = [&](int& t_, int& k_) {
      k_ += 2;
      while (k_ >= a.kblocks) { k_ -= a.kblocks; ++t_; }
    };
    float xc[16], zc[16], xn[16], zn[16];
    if (ti < my_tiles) gather(ti, kb, xc, zc);
    while (ti < my_tiles) {
      advance2(nti, nkb);
      const bool more = nti < my_tiles;
      const int cur_nt = rc.nt;          // n-tile of the CURRENT block (rc may move on in gather below)
      if (more) gather(nti, nkb, xn, zn);  // next block's loads fly while this one is processed
      mbar_wait(&empty[stage], sph ^ 1);
      if (r == 0) {
        const float* bsrc = a.Bp + ((size_t)cur_nt * a.kblocks + kb) * (size_t)(B_STAGE_BYTES / 4);
        mbar_arrive_expect_tx(&full[stage], B_STAGE_BYTES);
        bulk_g2s(Bs + stage * B_STAGE_BYTES, bsrc, B_STAGE_BYTES, &full[stage]);
      }
      uint8_t* abase = As + stage * A_STAGE_BYTES + r * 16;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        float h1[4], l1[4], h2[4], l2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xv = xc[kc * 4 + e];
          split_tf32(CONV ? xv : xv * zc[kc * 4 + e], h1[e], l1[e]);
          split_tf32(xv * xv, h2[e], l2[e]);
        }
        uint8_t* dst = abase + kc * (BM * 16);
        *reinterpret_cast<float4*>(dst + 0 * A_VARIANT_BYTES) = make_float4(h1[0], h1[1], h1[2], h1[3]);
        *reinterpret_cast<float4*>(dst + 1 * A_VARIANT_BYTES) = make_float4(l1[0], l1[1], l1[2], l1[3]);
        *reinterpret_cast<float4*>(dst + 2 * A_VARIANT_BYTES) = make_float4(h2[0], h2[1], h2[2], h2[3]);
        *reinterpret_cast<float4*>(dst + 3 * A_VARIANT_BYTES) = make_float4(l2[0], l2[1], l2[2], l2[3]);
      }
      fence_proxy_async();  // make this thread's st.shared visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stage]);
      if (more) {
#pragma unroll
        for (int e = 0; e < 16; ++e) { xc[e] = xn[e]; zc[e] = zn[e]; }
      }
      ti = nti; kb = nkb;
      stage += 2;
      if (stage >= STAGES) { stage -= STAGES; sph ^= 1; }
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

// ------------------------------------------------------------------------- weight packing
// Bp[nt][kb][v][kc][n][4]: v = {W hi, W lo, Vw hi, Vw lo}, zero padded in k and n; also vb.
template <int BN>
__global__ void pack_weights_kernel(int K, int N, int kblocks, int ntiles, const float* __restrict__ Wm,
                                    const float* __restrict__ Ls, float max_std, float* __restrict__ Bp,
                                    const float* __restrict__ lvb, float* __restrict__ vb) {
  const int64_t total = (int64_t)ntiles * kblocks * KC * BN * 4;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int e = (int)(idx & 3);
    int64_t t = idx >> 2;
    int n = (int)(t % BN);
    t /= BN;
    int kc = (int)(t % KC);
    t /= KC;
    int kb = (int)(t % kblocks);
    int nt = (int)(t / kblocks);
    int k = kb * BK + kc * 4 + e, gn = nt * BN + n;
    float w = 0.f, v = 0.f;
    if (k < K && gn < N) {
      w = Wm[(int64_t)k * N + gn];
      float s = fminf(expf(Ls[(int64_t)k * N + gn]), max_std);
      v = s * s;
    }
    float wh, wl, vh, vl;
    split_tf32(w, wh, wl);
    split_tf32(v, vh, vl);
    const int64_t stage_f = (int64_t)4 * KC * BN * 4;
    float* base = Bp + ((int64_t)nt * kblocks + kb) * stage_f + ((int64_t)kc * BN + n) * 4 + e;
    const int64_t vstride = (int64_t)KC * BN * 4;
    base[0 * vstride] = wh;
    base[1 * vstride] = wl;
    base[2 * vstride] = vh;
    base[3 * vstride] = vl;
    if (idx < N) vb[idx] = fminf(expf(lvb[idx]), max_std * max_std);
  }
}

template <int BN, bool CONV>
static int launch(mnf_ctx* ctx, Args& a, const float* mean_W, const float* log_std_W, const float* log_var_b,
                  float max_std) {
  a.kblocks = (a.K + BK - 1) / BK;
  a.ntiles = (a.N + BN - 1) / BN;
  const size_t bp_floats = (size_t)a.ntiles * a.kblocks * 4 * KC * BN * 4;
  bool fresh = true;
  float* ws = (float*)mnf_packed(ctx, mnf_mix(0x7C6Eull + BN + (CONV ? 1024 : 0), ((uint64_t)a.K << 32) | (uint32_t)a.N),
                                 (uint64_t)(uintptr_t)mean_W, (uint64_t)(uintptr_t)log_std_W,
                                 mnf_mix((uint64_t)(uintptr_t)log_var_b, (uint64_t)__float_as_uint_host(max_std)),
                                 sizeof(float) * (bp_floats + a.N + 64), &fresh);
  if (!ws) return MNF_ERR_CUDA;
  float* Bp = ws;
  float* vb = ws + bp_floats;
  if (fresh) {
    const int64_t tot = (int64_t)bp_floats / 4;
    int grid = (int)(cdiv(tot, 256) < ctx->sm_count * 8 ? cdiv(tot, 256) : ctx->sm_count * 8);
    pack_weights_kernel<BN><<<grid, 256, 0, ctx->stream>>>(a.K, a.N, a.kblocks, a.ntiles, mean_W, log_std_W, max_std, Bp,
                                                         log_var_b, vb);
    MNF_LAUNCHED(ctx);
  }
  a.Bp = Bp;
  a.vb = vb;
  const size_t tab_bytes = CONV ? sizeof(int) * 2 * (size_t)a.kblocks * BK / (a.vec ? 4 : 1) : 0;
  constexpr int STAGES = stages_for(BN);
  const size_t smem = (size_t)STAGES * (A_STAGE_BYTES + 4 * KC * BN * 16) +
                      sizeof(float) * ((direct_out(BN, CONV) ? 0 : BM * (BN + 1)) + 4) + 256 +
                      sizeof(float) * 2 * BN + tab_bytes;
  MNF_CHECK_ARG(smem <= 227 * 1024, "tensor-core conv path: K = %d is too large for the im2col table", a.K);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MNF_CUDA(cudaFuncSetAttribute(paired_umma_fwd<BN, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int64_t tiles = cdiv(a.M, BM) * a.ntiles;
  // as many CTAs as the busiest one needs rounds: 320 tiles on 148 SMs are 3 rounds either way, and
  // 107 CTAs x 3 tiles leave 41 SMs to the flow kernels running beside this one on the side streams
  const int64_t rounds = cdiv(tiles, ctx->sm_count);
  const int g = (int)cdiv(tiles, rounds);
  const bool probed = mnf_probe_begin(ctx, CONV ? MNF_PROBE_CONV_GEMM : MNF_PROBE_DENSE_GEMM);
  mnf_note_path(BN == 128 ? "paired_umma<128>" : "paired_umma<64>");
  paired_umma_fwd<BN, CONV><<<g, THREADS, smem, ctx->stream>>>(a);
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

}  // namespace tc

// Entry used by layers_fwd.cu when the context's math mode selects the tensor-core path.
// conv != 0: geo = {H,W,C,kh,kw,Ho,Wo,stride,pad_t,pad_l}.
int mnf_tc_paired_forward(mnf_ctx* ctx, int conv, int64_t M, int K, int N, const float* x, const float* z,
                          const float* mean_W, const float* log_std_W, const float* mean_b, const float* log_var_b,
                          float max_std, const NoiseDev& eps, float* y, float* mean_out, float* sd_out,
                          const int* geo, int act) {
  tc::Args a = {};
  a.pool = (conv && geo[10]) ? 1 : 0;
  a.act = conv ? 0 : act;
  a.M = M; a.K = K; a.N = N;
  a.x = x; a.z = z; a.bm = mean_b; a.y = y; a.mean_out = mean_out; a.sd_out = sd_out; a.eps = eps;
  if (conv) {
    a.H = geo[0]; a.W = geo[1]; a.C = geo[2]; a.kh = geo[3]; a.kw = geo[4]; a.Ho = geo[5]; a.Wo = geo[6];
    a.stride = geo[7]; a.pad_t = geo[8]; a.pad_l = geo[9];
    a.vec = (a.C % 4 == 0) && (((uintptr_t)x & 15) == 0);
    return tc::launch<64, true>(ctx, a, mean_W, log_std_W, log_var_b, max_std);
  }
  a.vec = (K % 4 == 0) && (((uintptr_t)x & 15) == 0) && (!z || ((uintptr_t)z & 15) == 0);
  // wide dense layers: 128-column tiles halve the number of times x and z are re-read from L2
  if (N >= 256 && !getenv("MNF_TC_BN64")) return tc::launch<128, false>(ctx, a, mean_W, log_std_W, log_var_b, max_std);
  return tc::launch<64, false>(ctx, a, mean_W, log_std_W, log_var_b, max_std);
}
