// IAF / RNVP flow step on the 5th-gen tensor cores (sm_100a): the whole coupling MLP of one
// step for a 128-row tile in ONE CTA, both layers as 3xTF32 tcgen05.mma with TMEM accumulators,
// the hidden activations never leaving the SM.
//   reference: IAF.forward = MAF.inverse maf.py:47-53,63 (MADE made.py:77-91, MaskedDense :28-30),
//              RNVP.forward rnvp.py:33-47, z0 sampling mnf_dense.py:93-96
//
//   stage 1   acc_h[128 x 64]  = in[128 x D] @ W1[D x H]      (in = z, or m*z for RNVP)
//             producers (warps 5-12) load/sample z, apply the Bernoulli mask, split hi/lo and
//             write K-major UMMA tiles; W1 arrives pre-split by cp.async.bulk; MMA warp 4.
//   stage 1e  epilogue warps 0-3: tcgen05.ld acc_h, + b1, ReLU/tanh, split hi/lo and write it
//             back to shared memory AS THE A OPERAND of stage 2 (never to HBM unless saved).
//   stage 2   per 32-column chunk: acc[128 x 64] = h @ [Ws chunk | Wt chunk]  (K = 64), weights
//             streamed by cp.async.bulk (warp 12), accumulators double-buffered in TMEM.
//   stage 2e  warps 0-3 and 8-11: (s,t) -> y = z*exp(s)+t (IAF, with reversal) or the gated
//             RNVP update; per-row log-det accumulated in registers, reduced through smem.
#include "common.cuh"

namespace fu {

constexpr int BM = 128;
constexpr int HP = 64;         // hidden units padded to 64
constexpr int BK = 16, KC = 4; // stage-1 k-block (floats) and its 16-byte chunks
constexpr int S1 = 4;          // stage-1 ring depth (S1 * 24 KB == S2 * 32 KB: the two rings alias)
constexpr int CN = 32;         // stage-2 columns per chunk (s and t: N = 64 per MMA)
constexpr int S2 = 3;          // stage-2 weight ring depth
constexpr int THREADS = 21 * 32;  // warps 0-3 stage-1 epilogue, 4 MMA, 5-12 and 13-20 producers; stage-2 epilogue: 0-3, 8-11, 13-20
constexpr int NG = 4;               // producer groups of 4 warps (128 rows), group g stages k-blocks g, g+NG, ...
constexpr int A1_VAR = KC * BM * 16;        // 8 KB: one variant (hi or lo) of a stage-1 A block
constexpr int A1_STAGE = 2 * A1_VAR;        // 16 KB
constexpr int B1_VAR = KC * HP * 16;        // 4 KB
constexpr int B1_STAGE = 2 * B1_VAR;        // 8 KB
constexpr int H_VAR = (HP / 4) * BM * 16;   // 32 KB: h hi (or lo) as a K = 64 A operand
constexpr int B2_VAR = (HP / 4) * 64 * 16;  // 16 KB: [K = 64][N = 64] weights, one variant
constexpr int B2_STAGE = 2 * B2_VAR;        // 32 KB
constexpr int TMEM_COLS = 256;              // acc_h: 64, stage-2 accumulators: 2 x 64

struct Args {
  int64_t B;
  int D, H, parity, kind;
  int kblocks;   // ceil(D / 16)
  int chunks;    // ceil(D / 32)
  const float* zin;
  const float* q0_mean;
  const float* q0_log_var;
  NoiseDev eps;
  float* z0_out;
  float* zout;
  const float* ld_in;
  float* ld_out;
  float* hid_save;
  const float* b1;
  const float* bs;
  const float* bt;
  const float* B1p;  // [kblocks][2][KC][HP][4]
  const float* B2p;  // [chunks][2][HP/4][64][4]
  NoiseDev bern;
};

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
  // Bounded (see umma.cuh): a barrier nobody will ever complete ends as a launch failure, not as a hung GPU.
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
// K-major SWIZZLE_NONE descriptor (see tc_gemm.cu): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1
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
// Warp-converged issue (see umma.cuh): all lanes execute, one elected lane performs the operation,
// so that warp-uniform descriptors stay in uniform registers instead of per-thread R2UR chains.
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
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// z (or the sampled z0) for elements [4*g4, 4*g4+4) of one row
__device__ __forceinline__ void load_z4(const Args& a, int64_t row, int g4, float out[4]) {
  const int D = a.D, d = g4 * 4;
  if (a.zin) {
    const float* p = a.zin + row * D + d;
    if (d + 3 < D && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p));
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
  for (int e = 0; e < 4; ++e) out[e] = d + e < D ? a.q0_mean[d + e] + expf(0.5f * a.q0_log_var[d + e]) * nrm[e] : 0.f;
}
// RNVP mask of the 16 elements of k-block kb of one row: bit e set = element kept.  Fair coins
// are one bit each (common.cuh: 128 per Philox block), so a thread re-runs Philox only when its
// k-blocks cross a 128-element boundary; `cache_blk` / `cache` carry the block between calls.
__device__ __forceinline__ uint32_t mask16(const NoiseDev& n, int64_t lrow, int kb, int n_inner, uint32_t& cache_blk,
                                           Philox4& cache) {
  if (n.injected) {
    uint32_t bits = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int d = kb * 16 + e;
      if (d < n_inner && n.injected[lrow * n_inner + d] != 0.f) bits |= 1u << e;
    }
    return bits;
  }
  const uint32_t blk = (uint32_t)kb >> 3;
  if (blk != cache_blk) {
    cache = philox_at(n, n.row_offset + (uint64_t)lrow, 0u, blk);
    cache_blk = blk;
  }
  const uint32_t w = philox_word(cache, ((uint32_t)kb >> 1) & 3u);
  return ~(w >> (((uint32_t)kb & 1u) * 16u)) & 0xFFFFu;  // draw = 1 when the bit is clear
}

__global__ void __launch_bounds__(THREADS, 1) flow_umma_step(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // region R: stage-1 rings, later re-used as the stage-2 weight ring
  uint8_t* A1 = smem;                         // [S1][2][KC][BM][16B]        64 KB
  uint8_t* B1 = A1 + S1 * A1_STAGE;           // [S1][2][KC][HP][16B]        32 KB
  uint8_t* B2 = smem;                         // [S2][2][16][64][16B]        96 KB (aliases A1/B1)
  uint8_t* Hs = smem + S2 * B2_STAGE;         // [2][16][BM][16B]            64 KB
  float* ldp = (float*)(Hs + 2 * H_VAR);      // [4][BM] log-det partials
  uint64_t* bars = (uint64_t*)(ldp + 4 * BM);
  uint64_t* full1 = bars;            // [S1]
  uint64_t* empty1 = full1 + S1;     // [S1]
  uint64_t* acc1_full = empty1 + S1; // [1]
  uint64_t* h_ready = acc1_full + 1; // [1]
  uint64_t* full2 = h_ready + 1;     // [S2]
  uint64_t* empty2 = full2 + S2;     // [S2]
  uint64_t* afull = empty2 + S2;     // [2]
  uint64_t* aempty = afull + 2;      // [2]
  uint32_t* tmem_slot = (uint32_t*)(aempty + 2);
  float* bias_s = (float*)(tmem_slot + 2);                       // [2][chunks*32]: bs | bt
  uint16_t* mbits = (uint16_t*)(bias_s + 2 * a.chunks * CN);     // [BM][kblocks] RNVP mask bits of each 16-col block

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int D = a.D;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S1; ++s) { mbar_init(&full1[s], 4 + 1); mbar_init(&empty1[s], 1); }
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t t_acc1 = tmem_base, t_acc2 = tmem_base + 64;  // acc2 buffers at +64, +128
  for (int c = threadIdx.x; c < a.chunks * CN; c += THREADS) {
    bias_s[c] = c < a.D ? a.bs[c] : 0.f;
    bias_s[a.chunks * CN + c] = c < a.D ? a.bt[c] : 0.f;
  }
  __syncthreads();

  // instruction descriptor: D = F32, A = B = TF32, K-major, M = 128, N = 64
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  // ---- stage-1 epilogue: h = act(acc1 + b1) -> hi/lo A operand of stage 2 in Hs.  Sixteen warps
  // share it (TMEM lane quadrant x 16 hidden columns): warps 0-3, and the producer warps 8-11,
  // 13-16, 17-20 once their k-blocks are published.  tanh by __expf (MUFU): ~1e-6 relative, inside
  // the parity bar like the stage-2 gate math.
  auto h_epilogue = [&](int quad, int c0) {
    const int r = quad * 32 + lane;
    const int64_t row = row0 + r;
    mbar_wait(acc1_full, 0);
    tc_fence_after();
    const uint32_t trow = t_acc1 + ((uint32_t)(quad * 32) << 16);
    float v[16];
    tmem_ld16(trow + c0, v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float h4[4], l4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int h = c0 + 4 * q + e;
        float hv = 0.f;
        if (h < a.H) {
          const float pre = v[4 * q + e] + __ldg(a.b1 + h);
          if (a.kind == MNF_FLOW_IAF) hv = fmaxf(pre, 0.f);
          else {
            const float t = __expf(-2.f * fabsf(pre));         // in (0, 1]
            const float th = __fdividef(1.f - t, 1.f + t);     // tanh(|pre|)
            hv = pre < 0.f ? -th : th;
          }
          if (a.hid_save && row < a.B) a.hid_save[row * a.H + h] = hv;
        }
        split_tf32(hv, h4[e], l4[e]);
      }
      const int kc = (c0 >> 2) + q;  // 16-byte chunk index along K = hidden
      *reinterpret_cast<float4*>(Hs + kc * (BM * 16) + r * 16) = make_float4(h4[0], h4[1], h4[2], h4[3]);
      *reinterpret_cast<float4*>(Hs + H_VAR + kc * (BM * 16) + r * 16) = make_float4(l4[0], l4[1], l4[2], l4[3]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(h_ready);
  };

  if (warp >= 5) {
    // =========================================================== stage-1 producers (warps 5-20)
    const int pt = threadIdx.x - 5 * 32;   // 0..511: four groups of 128 rows
    const int r = pt & (BM - 1), grp = pt >> 7;
    int64_t row = row0 + r;
    const bool row_ok = row < a.B;
    if (!row_ok) row = a.B - 1;
    float cur[16], nxt[16];
    uint32_t cblk = 0xFFFFFFFFu;  // fair-coin cache (mask16)
    Philox4 cq = {0u, 0u, 0u, 0u};
    auto gather = [&](int kb, float v[16]) {
      uint32_t bits = 0;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        const int g4 = kb * KC + kc;
        if (g4 * 4 < D) {
          load_z4(a, row, g4, &v[kc * 4]);
          if (a.z0_out && !a.zin && row_ok) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (g4 * 4 + e < D) a.z0_out[row * D + g4 * 4 + e] = v[kc * 4 + e];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[kc * 4 + e] = 0.f;
        }
      }
      if (a.kind == MNF_FLOW_RNVP) {
        bits = mask16(a.bern, row, kb, D, cblk, cq);
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (!((bits >> e) & 1u)) v[e] = 0.f;
      }
      if (a.kind == MNF_FLOW_RNVP) mbits[r * a.kblocks + kb] = (uint16_t)bits;  // re-used by the stage-2 epilogue
    };
    // split one k-block (16 values of this row) into the UMMA A tiles of its ring stage
    auto publish = [&](int kb, const float v[16]) {
      const int stage = kb % S1;
      const uint32_t sph = (uint32_t)((kb / S1) & 1);
      mbar_wait(&empty1[stage], sph ^ 1);
      if (r == 0) {
        mbar_arrive_expect_tx(&full1[stage], B1_STAGE);
        bulk_g2s(B1 + stage * B1_STAGE, a.B1p + (size_t)kb * (B1_STAGE / 4), B1_STAGE, &full1[stage]);
      }
      uint8_t* dst0 = A1 + stage * A1_STAGE + r * 16;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32(v[kc * 4 + e], h[e], l[e]);
        *reinterpret_cast<float4*>(dst0 + kc * (BM * 16)) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(dst0 + A1_VAR + kc * (BM * 16)) = make_float4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full1[stage]);
    };
    const bool use_async = a.zin && (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.zin) & 15) == 0);
    if (use_async) {
      // deep prefetch: this thread's 64 bytes of each k-block travel global -> shared by cp.async
      // into a raw ring (8 blocks in flight per CTA; the ring borrows the h buffers, which are
      // not written before stage 1 has drained).  A thread only ever reads back its own bytes.
      constexpr int RAW = 8, PF = RAW / NG;
      const float* zrow = a.zin + row * D;
      const int sw = (r >> 1) & 3;
      auto issue = [&](int kb) {
        const uint32_t dst = smem_u32(Hs + (kb % RAW) * (BM * 64) + r * 64);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          const int d = (kb * KC + kc) * 4;
          if (d < D)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + ((kc ^ sw) << 4)), "l"(zrow + d) : "memory");
        }
      };
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        if (grp + NG * i < a.kblocks) issue(grp + NG * i);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      for (int kb = grp; kb < a.kblocks; kb += NG) {
        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");
        const uint8_t* src = Hs + (kb % RAW) * (BM * 64) + r * 64;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((kb * KC + kc) * 4 < D) v = *reinterpret_cast<const float4*>(src + ((kc ^ sw) << 4));
          cur[kc * 4 + 0] = v.x; cur[kc * 4 + 1] = v.y; cur[kc * 4 + 2] = v.z; cur[kc * 4 + 3] = v.w;
        }
        if (kb + RAW < a.kblocks) issue(kb + RAW);  // same slot, already copied to registers
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (a.kind == MNF_FLOW_RNVP) {
          const uint32_t bits = mask16(a.bern, row, kb, D, cblk, cq);
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (!((bits >> e) & 1u)) cur[e] = 0.f;
          mbits[r * a.kblocks + kb] = (uint16_t)bits;
        }
        publish(kb, cur);
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
      if (grp < a.kblocks) gather(grp, cur);
      for (int kb = grp; kb < a.kblocks; kb += NG) {
        const bool more = kb + NG < a.kblocks;
        if (more) gather(kb + NG, nxt);
        publish(kb, cur);
        if (more) {
#pragma unroll
          for (int e = 0; e < 16; ++e) cur[e] = nxt[e];
        }
      }
    }
    // ---- the producer warps that also serve stage 2 take their share of the stage-1 epilogue
    if (warp >= 8 && warp < 12) h_epilogue(warp & 3, 16);
    else if (warp >= 13 && warp < 17) h_epilogue(warp & 3, 32);
    else if (warp >= 17) h_epilogue(warp & 3, 48);
    // ---- warp 12: stage-2 weight loader (the ring re-uses the stage-1 buffers, so wait until
    //      stage 1 has fully drained: acc1_full fires after the last stage-1 MMA retired)
    if (warp == 12) {
      mbar_wait(acc1_full, 0);
      if (lane == 0) {
        int s2 = 0;
        uint32_t p2 = 0;
        for (int c = 0; c < a.chunks; ++c) {
          mbar_wait(&empty2[s2], p2 ^ 1);
          mbar_arrive_expect_tx(&full2[s2], B2_STAGE);
          bulk_g2s(B2 + s2 * B2_STAGE, a.B2p + (size_t)c * (B2_STAGE / 4), B2_STAGE, &full2[s2]);
          if (++s2 == S2) { s2 = 0; p2 ^= 1; }
        }
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // ================================================================================ MMA warp
    const uint64_t a1d = make_desc(smem_u32(A1), BM * 16, 128), b1d = make_desc(smem_u32(B1), HP * 16, 128);
    int stage = 0;
    uint32_t sph = 0;
    for (int kb = 0; kb < a.kblocks; ++kb) {
      mbar_wait(&full1[stage], sph);
      tc_fence_after();
      {
        const uint64_t ad = a1d + (uint64_t)((stage * A1_STAGE) >> 4), bd = b1d + (uint64_t)((stage * B1_STAGE) >> 4);
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          const uint64_t ah = ad + (uint64_t)((2 * j * BM * 16) >> 4), al = ah + (A1_VAR >> 4);
          const uint64_t bh = bd + (uint64_t)((2 * j * HP * 16) >> 4), bl = bh + (B1_VAR >> 4);
          umma_tf32_e(t_acc1, ah, bh, idesc, (kb == 0 && j == 0) ? 0u : 1u);
          umma_tf32_e(t_acc1, al, bh, idesc, 1u);
          umma_tf32_e(t_acc1, ah, bl, idesc, 1u);
        }
        umma_commit_e(&empty1[stage]);
        if (kb == a.kblocks - 1) umma_commit_e(acc1_full);
      }
      __syncwarp();
      if (++stage == S1) { stage = 0; sph ^= 1; }
    }
    // stage 2: h (hi/lo in Hs) @ weight chunks
    mbar_wait(h_ready, 0);
    tc_fence_after();
    const uint64_t hd = make_desc(smem_u32(Hs), BM * 16, 128), b2d = make_desc(smem_u32(B2), 64 * 16, 128);
    int s2 = 0;
    uint32_t p2 = 0;
    for (int c = 0; c < a.chunks; ++c) {
      const int buf = c & 1;
      mbar_wait(&aempty[buf], ((c >> 1) & 1) ^ 1);
      mbar_wait(&full2[s2], p2);
      tc_fence_after();
      {
        const uint64_t bd = b2d + (uint64_t)((s2 * B2_STAGE) >> 4);
        const uint32_t dacc = t_acc2 + buf * 64;
#pragma unroll
        for (int j = 0; j < HP / 8; ++j) {
          const uint64_t ah = hd + (uint64_t)((2 * j * BM * 16) >> 4), al = ah + (H_VAR >> 4);
          const uint64_t bh = bd + (uint64_t)((2 * j * 64 * 16) >> 4), bl = bh + (B2_VAR >> 4);
          umma_tf32_e(dacc, ah, bh, idesc, j == 0 ? 0u : 1u);
          umma_tf32_e(dacc, al, bh, idesc, 1u);
          umma_tf32_e(dacc, ah, bl, idesc, 1u);
        }
        umma_commit_e(&empty2[s2]);
        umma_commit_e(&afull[buf]);
      }
      __syncwarp();
      if (++s2 == S2) { s2 = 0; p2 ^= 1; }
    }
  } else {
    // ======================================================= stage-1 epilogue, hidden columns 0-15
    h_epilogue(warp, 0);
  }

  // ============================================ stage-2 epilogue: warps 0-3 (cols 0-15) and 8-11 (cols 16-31)
  // Fast-math note: exp / log / reciprocal use the MUFU approximations (__expf, __logf,
  // __fdividef): relative error ~1e-6 for the |s| this path sees, inside the 1e-5 parity bar
  // (tests/test_gpu_tensorcore.py); the FFMA kernels keep the IEEE-accurate versions.
  if (warp < 4 || (warp >= 8 && warp < 12) || warp >= 13) {
    // 16 epilogue warps: TMEM lane quadrant g x column half x chunk parity (= accumulator buffer)
    const int g = warp & 3;
    const int half = (warp >= 8 && warp < 12) || warp >= 17 ? 1 : 0, par = warp >= 13 ? 1 : 0;
    const int r = g * 32 + lane;
    const int64_t row = row0 + r;
    const bool row_ok = row < a.B;
    const float* zsrc = a.zin ? a.zin : a.z0_out;  // the sampled z0 was stored by the producers
    const float* zrow = zsrc + (row_ok ? row : 0) * D;
    const bool vec_ok = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(zsrc) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(a.zout) & 15) == 0);
    const float* bs_s = bias_s;
    const float* bt_s = bias_s + a.chunks * CN;
    float ld = 0.f;
    for (int c = par; c < a.chunks; c += 2) {
      const int buf = c & 1;
      const int cb = c * CN + half * 16;
      // z of this thread's 16 columns: issued before the accumulator wait so the latency overlaps
      float z16[16];
      if (a.zin) {  // (a sampled z0 only becomes visible after the barrier chain below)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c4 = cb + 4 * q;
          if (row_ok && vec_ok && c4 + 3 < D) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(zrow + c4));
            z16[4 * q + 0] = v.x; z16[4 * q + 1] = v.y; z16[4 * q + 2] = v.z; z16[4 * q + 3] = v.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) z16[4 * q + e] = (row_ok && c4 + e < D) ? __ldg(zrow + c4 + e) : 0.f;
          }
        }
      }
      mbar_wait(&afull[buf], (c >> 1) & 1);
      tc_fence_after();
      const uint32_t tb = t_acc2 + buf * 64 + ((uint32_t)(g * 32) << 16) + half * 16;
      float sv[16], tv[16];
      tmem_ld16(tb, sv);
      tmem_ld16(tb + 32, tv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aempty[buf]);
      if (!a.zin) {
#pragma unroll
        for (int e = 0; e < 16; ++e) z16[e] = (row_ok && cb + e < D) ? zrow[cb + e] : 0.f;
      }
      if (row_ok && cb < D) {
        const uint32_t mb = (a.kind == MNF_FLOW_RNVP) ? (uint32_t)mbits[r * a.kblocks + (cb >> 4)] : 0u;
        float out[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float s = sv[e] + bs_s[cb + e], t = tv[e] + bt_s[cb + e], zv = z16[e];
          if (a.kind == MNF_FLOW_IAF) {
            out[e] = fmaf(zv, __expf(s), t);
            if (cb + e < D) ld += s;
          } else {
            const float m = (mb >> e) & 1u ? 1.f : 0.f;
            const float u = __expf(-fabsf(s));           // in (0, 1]
            const float rcp = __fdividef(1.f, 1.f + u);
            const float gate = s >= 0.f ? rcp : u * rcp;  // sigmoid(s)
            const float lg = fminf(s, 0.f) - __logf(1.f + u);  // log(sigmoid(s))
            out[e] = ((1.f - m) * zv * gate + (1.f - gate) * t) + m * zv;
            if (cb + e < D) ld += (1.f - m) * lg;
          }
        }
        if (a.kind == MNF_FLOW_IAF && a.parity) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (cb + e < D) a.zout[row * D + (D - 1 - cb - e)] = out[e];
        } else {
          float* op = a.zout + row * D + cb;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (vec_ok && cb + 4 * q + 3 < D) {
              *reinterpret_cast<float4*>(op + 4 * q) = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (cb + 4 * q + e < D) op[4 * q + e] = out[4 * q + e];
            }
          }
        }
      }
    }
    ldp[(half + 2 * par) * BM + r] = ld;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < BM) {
    const int64_t row = row0 + threadIdx.x;
    if (row < a.B && a.ld_out)
      a.ld_out[row] = (a.ld_in ? a.ld_in[row] : 0.f) + ((ldp[threadIdx.x] + ldp[BM + threadIdx.x]) + (ldp[2 * BM + threadIdx.x] + ldp[3 * BM + threadIdx.x]));
  }
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// W1 [D,H] (masked) -> B1p[kb][v][kc][n = hidden][4 k];  [Ws|Wt] -> B2p[chunk][v][kc (hidden/4)][n][4 hidden]
__global__ void pack_flow_weights(int D, int H, int kblocks, int chunks, const float* __restrict__ w1,
                                  const float* __restrict__ m1, const float* __restrict__ ws,
                                  const float* __restrict__ wt, const float* __restrict__ m2, int64_t ld2,
                                  float* __restrict__ B1p, float* __restrict__ B2p) {
  const int64_t n1 = (int64_t)kblocks * KC * HP * 4;
  const int64_t n2 = (int64_t)chunks * (HP / 4) * 64 * 4;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n1 + n2; idx += (int64_t)gridDim.x * blockDim.x) {
    float w = 0.f;
    float* dst;
    int64_t vstride;
    if (idx < n1) {
      const int e = (int)(idx & 3);
      int64_t t = idx >> 2;
      const int n = (int)(t % HP);
      t /= HP;
      const int kc = (int)(t % KC);
      const int kb = (int)(t / KC);
      const int k = kb * BK + kc * 4 + e;
      if (k < D && n < H) {
        w = w1[(int64_t)k * H + n];
        if (m1) w *= m1[(int64_t)k * H + n];
      }
      dst = B1p + (int64_t)kb * (B1_STAGE / 4) + ((int64_t)kc * HP + n) * 4 + e;
      vstride = B1_VAR / 4;
    } else {
      const int64_t i2 = idx - n1;
      const int e = (int)(i2 & 3);
      int64_t t = i2 >> 2;
      const int n = (int)(t % 64);
      t /= 64;
      const int kc = (int)(t % (HP / 4));
      const int c = (int)(t / (HP / 4));
      const int h = kc * 4 + e, col = c * CN + (n & 31);
      if (h < H && col < D) {
        const int64_t o = (int64_t)h * ld2 + col;
        w = n < 32 ? ws[o] : wt[o];
        if (m2) w *= n < 32 ? m2[o] : m2[o + D];
      }
      dst = B2p + (int64_t)c * (B2_STAGE / 4) + ((int64_t)kc * 64 + n) * 4 + e;
      vstride = B2_VAR / 4;
    }
    float hi, lo;
    split_tf32(w, hi, lo);
    dst[0] = hi;
    dst[vstride] = lo;
  }
}

}  // namespace fu

// One IAF/RNVP step on tensor cores; ws_off: byte offset into the context workspace to use.
int mnf_flow_umma_step(mnf_ctx* ctx, int64_t B, int D, int H, int kind, int parity, const float* zin,
                       const float* q0_mean, const float* q0_log_var, const NoiseDev& eps, float* z0_out, float* zout,
                       const float* ld_in, float* ld_out, float* hid_save, const mnf_flow_step& st,
                       const NoiseDev& bern, float* pack_ws) {
  using namespace fu;
  Args a = {};
  a.B = B; a.D = D; a.H = H; a.parity = parity; a.kind = kind;
  a.kblocks = (D + BK - 1) / BK;
  a.chunks = (D + CN - 1) / CN;
  a.zin = zin; a.q0_mean = q0_mean; a.q0_log_var = q0_log_var; a.eps = eps; a.z0_out = z0_out; a.zout = zout;
  a.ld_in = ld_in; a.ld_out = ld_out; a.hid_save = hid_save;
  a.b1 = st.b1; a.bs = st.bs; a.bt = st.bt; a.bern = bern;
  bool fresh = true;
  if (ctx->frozen) {  // frozen weights: one persistent packed block per flow step
    pack_ws = (float*)mnf_packed(ctx, mnf_mix(0xF10Full, ((uint64_t)D << 32) | (uint32_t)H), (uint64_t)(uintptr_t)st.w1,
                                 mnf_mix((uint64_t)(uintptr_t)st.ws, (uint64_t)(uintptr_t)st.wt),
                                 mnf_mix((uint64_t)(uintptr_t)st.m1, mnf_mix((uint64_t)(uintptr_t)st.m2, (uint64_t)st.ld2)),
                                 sizeof(float) * mnf_flow_umma_pack_floats(D), &fresh);
    if (!pack_ws) return MNF_ERR_CUDA;
  }
  float* B1p = pack_ws;
  float* B2p = pack_ws + (size_t)a.kblocks * (B1_STAGE / 4);
  a.B1p = B1p; a.B2p = B2p;
  if (fresh) {
    const int64_t tot = (int64_t)a.kblocks * KC * HP * 4 + (int64_t)a.chunks * (HP / 4) * 64 * 4;
    int grid = (int)(cdiv(tot, 256) < ctx->sm_count * 4 ? cdiv(tot, 256) : ctx->sm_count * 4);
    pack_flow_weights<<<grid, 256, 0, ctx->stream>>>(D, H, a.kblocks, a.chunks, st.w1, st.m1, st.ws, st.wt, st.m2, st.ld2,
                                                    B1p, B2p);
    MNF_LAUNCHED(ctx);
  }
  const size_t smem = (size_t)S2 * B2_STAGE + 2 * H_VAR + sizeof(float) * 4 * BM + 256 +
                      sizeof(float) * 2 * (size_t)a.chunks * CN +
                      (kind == MNF_FLOW_RNVP ? sizeof(uint16_t) * (size_t)BM * a.kblocks : 0) + 16;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MNF_CUDA(cudaFuncSetAttribute(flow_umma_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const bool probed = mnf_probe_begin(ctx, MNF_PROBE_FLOW_STEP);
  mnf_note_path("flow_umma");
  flow_umma_step<<<(unsigned)cdiv(B, BM), THREADS, smem, ctx->stream>>>(a);
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}

size_t mnf_flow_umma_pack_floats(int D) {
  using namespace fu;
  const size_t kblocks = (D + BK - 1) / BK, chunks = (D + CN - 1) / CN;
  return kblocks * (B1_STAGE / 4) + chunks * (B2_STAGE / 4) + 64;
}
