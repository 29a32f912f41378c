// MNF_MATH_BF16: the paired mean / variance contraction of MNFDense.call (mnf_dense.py:159-170)
// with bf16 operands and fp32 accumulation -- BASELINE.json's "wide MNFDense 4096 -> 4096, batch
// 8192, bf16 tensor-core path" (SURVEY.md section 8d, C4).  One tcgen05.mma.kind::f16 per product
// and K = 16 step (no hi/lo split: 8 significant bits per operand, tolerance stated in
// tests/test_gpu_bf16.py), accumulators in TMEM.
//
// Both operand streams are packed ONCE per call into the no-swizzle K-major UMMA layout, so the
// persistent GEMM kernel's producer is two cp.async.bulk (TMA, 1-D) per 64-wide k-stage:
//   pack_rows_bf16    x, z -> Ap[m-tile][k-stage][{x*z, x^2}][8 chunks][128 rows][8 bf16]   (32 KB)
//   pack_weights_bf16 W, log sigma -> Bp[n-tile][k-stage][{W, clip(exp(L))^2}][8][128][8]    (32 KB)
// x is read once for both products; z-scaling and squaring happen in fp32 before the rounding.
//   paired_bf16_fwd   warps 0-3 epilogue (tcgen05.ld, bias, sqrt, Philox eps, ReLU, rows stored
//                     from registers), warp 4 MMA issue (warp-converged, elect.sync), warp 5 TMA
//                     producer; three 64 KB stages; TMEM 512 columns = 2 buffers x {mean, var}.
#include <cuda_bf16.h>

#include "umma.cuh"

namespace db {
using namespace um;

constexpr int BM = 128, BN = 128;
constexpr int KS = 64;                       // k per pipeline stage
constexpr int CH = KS / 8;                   // 16-byte chunks (8 bf16) per operand row and stage
constexpr int A_VAR_BYTES = CH * BM * 16;    // 16 KB
constexpr int A_STAGE_BYTES = 2 * A_VAR_BYTES;
constexpr int B_VAR_BYTES = CH * BN * 16;
constexpr int B_STAGE_BYTES = 2 * B_VAR_BYTES;
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 4;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int TMEM_COLS = 4 * BN;            // 2 buffers x {mean, var}

struct Args {
  int64_t M, mtiles;
  int K, N, kblocks, ntiles;
  const uint8_t* Ap;
  const uint8_t* Bp;
  const float* bm;
  const float* vb;
  float* y;
  float* sd_out;
  NoiseDev eps;
  int act;  // 1 = ReLU fused into the epilogue
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// one thread = one 16-byte chunk (8 consecutive k of one row) of both variants
__global__ void pack_rows_bf16(int64_t M, int K, int kblocks, int64_t mtiles, const float* __restrict__ x,
                               const float* __restrict__ z, uint8_t* __restrict__ Ap) {
  const int64_t total = mtiles * kblocks * (int64_t)(CH * BM);
  const bool vec = (K % 4 == 0) && (((uintptr_t)x & 15) == 0) && (!z || ((uintptr_t)z & 15) == 0);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    // consecutive threads walk the k chunks of a row first (256 contiguous bytes of x), then rows
    const int c = (int)(idx % CH);
    int64_t t = idx / CH;
    const int r = (int)(t % BM);
    t /= BM;
    const int kb = (int)(t % kblocks);
    const int64_t mt = t / kblocks;
    const int64_t m = mt * BM + r;
    const int k0 = kb * KS + c * 8;
    float xv[8], zv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { xv[e] = 0.f; zv[e] = 1.f; }
    if (m < M) {
      const float* xr = x + m * K + k0;
      const float* zr = z ? z + m * K + k0 : nullptr;
      if (vec && k0 + 8 <= K) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(xr)), a1 = __ldg(reinterpret_cast<const float4*>(xr) + 1);
        xv[0] = a0.x; xv[1] = a0.y; xv[2] = a0.z; xv[3] = a0.w; xv[4] = a1.x; xv[5] = a1.y; xv[6] = a1.z; xv[7] = a1.w;
        if (zr) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(zr)), b1 = __ldg(reinterpret_cast<const float4*>(zr) + 1);
          zv[0] = b0.x; zv[1] = b0.y; zv[2] = b0.z; zv[3] = b0.w; zv[4] = b1.x; zv[5] = b1.y; zv[6] = b1.z; zv[7] = b1.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (k0 + e < K) {
            xv[e] = __ldg(xr + e);
            if (zr) zv[e] = __ldg(zr + e);
          }
      }
    }
    uint4 am, av;
    am.x = pack_bf16x2(xv[0] * zv[0], xv[1] * zv[1]); am.y = pack_bf16x2(xv[2] * zv[2], xv[3] * zv[3]);
    am.z = pack_bf16x2(xv[4] * zv[4], xv[5] * zv[5]); am.w = pack_bf16x2(xv[6] * zv[6], xv[7] * zv[7]);
    av.x = pack_bf16x2(xv[0] * xv[0], xv[1] * xv[1]); av.y = pack_bf16x2(xv[2] * xv[2], xv[3] * xv[3]);
    av.z = pack_bf16x2(xv[4] * xv[4], xv[5] * xv[5]); av.w = pack_bf16x2(xv[6] * xv[6], xv[7] * xv[7]);
    uint8_t* dst = Ap + (mt * kblocks + kb) * (int64_t)A_STAGE_BYTES + ((int64_t)c * BM + r) * 16;
    *reinterpret_cast<uint4*>(dst) = am;
    *reinterpret_cast<uint4*>(dst + A_VAR_BYTES) = av;
  }
}

// one thread = one 16-byte chunk (8 consecutive k of one output column) of both variants; also vb
__global__ void pack_weights_bf16(int K, int N, int kblocks, int ntiles, const float* __restrict__ Wm,
                                  const float* __restrict__ Ls, float max_std, uint8_t* __restrict__ Bp,
                                  const float* __restrict__ lvb, float* __restrict__ vb) {
  const int64_t total = (int64_t)ntiles * kblocks * (CH * BN);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    if (idx < N) vb[idx] = fminf(expf(lvb[idx]), max_std * max_std);
    const int n = (int)(idx % BN);  // consecutive threads read consecutive columns of a weight row
    int64_t t = idx / BN;
    const int c = (int)(t % CH);
    t /= CH;
    const int kb = (int)(t % kblocks);
    const int nt = (int)(t / kblocks);
    const int gn = nt * BN + n, k0 = kb * KS + c * 8;
    float w[8], v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      w[e] = v[e] = 0.f;
      if (gn < N && k0 + e < K) {
        w[e] = __ldg(Wm + (int64_t)(k0 + e) * N + gn);
        const float s = fminf(expf(__ldg(Ls + (int64_t)(k0 + e) * N + gn)), max_std);
        v[e] = s * s;
      }
    }
    uint4 bw, bv;
    bw.x = pack_bf16x2(w[0], w[1]); bw.y = pack_bf16x2(w[2], w[3]); bw.z = pack_bf16x2(w[4], w[5]); bw.w = pack_bf16x2(w[6], w[7]);
    bv.x = pack_bf16x2(v[0], v[1]); bv.y = pack_bf16x2(v[2], v[3]); bv.z = pack_bf16x2(v[4], v[5]); bv.w = pack_bf16x2(v[6], v[7]);
    uint8_t* dst = Bp + ((int64_t)nt * kblocks + kb) * (int64_t)B_STAGE_BYTES + ((int64_t)c * BN + n) * 16;
    *reinterpret_cast<uint4*>(dst) = bw;
    *reinterpret_cast<uint4*>(dst + B_VAR_BYTES) = bv;
  }
}

__global__ void __launch_bounds__(THREADS, 1) paired_bf16_fwd(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* As = smem;                          // [STAGES][2][CH][BM][16 B]
  uint8_t* Bs = As + STAGES * A_STAGE_BYTES;   // [STAGES][2][CH][BN][16 B]
  uint64_t* full = (uint64_t*)(Bs + STAGES * B_STAGE_BYTES);
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

  if (warp < EPI_WARPS) {
    // ================================================================== epilogue warps
    const int r = warp * 32 + lane;  // tile row == TMEM lane
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntot; t += gridDim.x, ++it) {
      const int64_t mt = t / a.ntiles;
      const int nt = (int)(t - mt * a.ntiles);
      const int64_t m = mt * BM + r;
      const int n0 = nt * BN;
      const int buf = it & 1;
      mbar_wait(&tfull[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const bool row_ok = m < a.M;
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
              philox_normal4(a.eps, a.eps.row_offset + (uint64_t)m, 0u, (uint32_t)(nb >> 2), e);
            }
            float o4[4] = {0.f, 0.f, 0.f, 0.f}, s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = nb + j;
              if (n < a.N) {
                const float mean = mv[4 * q + j] + __ldg(a.bm + n);
                const float sd = sqrtf(vv[4 * q + j] + __ldg(a.vb + n));
                s4[j] = sd;
                const float out = fmaf(sd, e[j], mean);
                o4[j] = a.act == 1 ? fmaxf(out, 0.f) : out;
              }
            }
            float* dstp = a.y + m * a.N + nb;
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
      // accumulators consumed: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
  } else if (warp == EPI_WARPS) {
    // ======================================================================= MMA issuer
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, K-major both
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint64_t adesc0 = make_desc(smem_u32(As), BM * 16, 128), bdesc0 = make_desc(smem_u32(Bs), BN * 16, 128);
    int it = 0, stage = 0;
    uint32_t sph = 0;
    for (int64_t t = blockIdx.x; t < ntot; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_mean = tmem_base + (uint32_t)(buf * 2 * BN), d_var = d_mean + BN;
      for (int kb = 0; kb < a.kblocks; ++kb) {
        mbar_wait(&full[stage], sph);
        tc_fence_after();
        const uint64_t ad = adesc0 + (uint64_t)((stage * A_STAGE_BYTES) >> 4);
        const uint64_t bd = bdesc0 + (uint64_t)((stage * B_STAGE_BYTES) >> 4);
#pragma unroll
        for (int j = 0; j < KS / 16; ++j) {  // one K = 16 instruction = two 16-byte chunks
          const uint64_t ao = ad + (uint64_t)((2 * j * BM * 16) >> 4), bo = bd + (uint64_t)((2 * j * BN * 16) >> 4);
          const uint32_t acc = (kb == 0 && j == 0) ? 0u : 1u;
          umma_f16_elect(d_mean, ao, bo, idesc, acc);
          umma_f16_elect(d_var, ao + (A_VAR_BYTES >> 4), bo + (B_VAR_BYTES >> 4), idesc, acc);
        }
        umma_commit_elect(&empty[stage]);                         // smem stage free once these MMAs retire
        if (kb == a.kblocks - 1) umma_commit_elect(&tfull[buf]);  // accumulators ready
        __syncwarp();
        if (++stage == STAGES) { stage = 0; sph ^= 1; }
      }
    }
  } else {
    // ======================================================================== TMA producer
    int stage = 0;
    uint32_t sph = 0;
    for (int64_t t = blockIdx.x; t < ntot; t += gridDim.x) {
      const int64_t mt = t / a.ntiles;
      const int nt = (int)(t - mt * a.ntiles);
      const uint8_t* asrc = a.Ap + mt * a.kblocks * (int64_t)A_STAGE_BYTES;
      const uint8_t* bsrc = a.Bp + (int64_t)nt * a.kblocks * (int64_t)B_STAGE_BYTES;
      for (int kb = 0; kb < a.kblocks; ++kb) {
        mbar_wait(&empty[stage], sph ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full[stage], A_STAGE_BYTES + B_STAGE_BYTES);
          bulk_g2s(As + stage * A_STAGE_BYTES, asrc + (int64_t)kb * A_STAGE_BYTES, A_STAGE_BYTES, &full[stage]);
          bulk_g2s(Bs + stage * B_STAGE_BYTES, bsrc + (int64_t)kb * B_STAGE_BYTES, B_STAGE_BYTES, &full[stage]);
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

}  // namespace db

// Entry used by layers_fwd.cu when the context's math mode is MNF_MATH_BF16.
int mnf_dense_bf16_forward(mnf_ctx* ctx, int64_t M, int K, int N, const float* x, const float* z, const float* mean_W,
                           const float* log_std_W, const float* mean_b, const float* log_var_b, float max_std,
                           const NoiseDev& eps, float* y, float* sd_out, int act) {
  using namespace db;
  Args a = {};
  a.M = M; a.K = K; a.N = N;
  a.mtiles = cdiv(M, BM);
  a.kblocks = (K + KS - 1) / KS;
  a.ntiles = (N + BN - 1) / BN;
  a.bm = mean_b; a.y = y; a.sd_out = sd_out; a.eps = eps; a.act = act;
  const size_t a_bytes = (size_t)a.mtiles * a.kblocks * A_STAGE_BYTES;
  const size_t b_bytes = (size_t)a.ntiles * a.kblocks * B_STAGE_BYTES + sizeof(float) * (((size_t)N + 63) & ~(size_t)63);
  uint8_t *Ap, *Bp;
  bool fresh = true;
  if (ctx->frozen) {
    // the packed weights persist (keyed by the parameter buffers); only the rows are packed per call
    Bp = (uint8_t*)mnf_packed(ctx, mnf_mix(0xBF16ull, ((uint64_t)K << 32) | (uint32_t)N), (uint64_t)(uintptr_t)mean_W,
                              (uint64_t)(uintptr_t)log_std_W,
                              mnf_mix((uint64_t)(uintptr_t)log_var_b, (uint64_t)__float_as_uint_host(max_std)), b_bytes, &fresh);
    if (!Bp) return MNF_ERR_CUDA;
    Ap = (uint8_t*)mnf_workspace(ctx, a_bytes);
    if (!Ap) return MNF_ERR_CUDA;
  } else {
    Ap = (uint8_t*)mnf_workspace(ctx, a_bytes + b_bytes);
    if (!Ap) return MNF_ERR_CUDA;
    Bp = Ap + a_bytes;
  }
  float* vb = (float*)(Bp + (size_t)a.ntiles * a.kblocks * B_STAGE_BYTES);
  if (fresh) {
    const int64_t tot = (int64_t)a.ntiles * a.kblocks * (CH * BN);
    const int grid = (int)(cdiv(tot, 256) < (int64_t)ctx->sm_count * 16 ? cdiv(tot, 256) : (int64_t)ctx->sm_count * 16);
    pack_weights_bf16<<<grid, 256, 0, ctx->stream>>>(K, N, a.kblocks, a.ntiles, mean_W, log_std_W, max_std, Bp, log_var_b, vb);
    MNF_LAUNCHED(ctx);
  }
  {
    const int64_t tot = a.mtiles * a.kblocks * (int64_t)(CH * BM);
    const int grid = (int)(cdiv(tot, 256) < (int64_t)ctx->sm_count * 16 ? cdiv(tot, 256) : (int64_t)ctx->sm_count * 16);
    pack_rows_bf16<<<grid, 256, 0, ctx->stream>>>(M, K, a.kblocks, a.mtiles, x, z, Ap);
    MNF_LAUNCHED(ctx);
  }
  a.Ap = Ap; a.Bp = Bp; a.vb = vb;
  const size_t smem = (size_t)STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256;
  static bool attr_set[64] = {};  // per device: function attributes belong to the device's copy of the kernel
  if (!attr_set[ctx->device & 63]) {
    MNF_CUDA(cudaFuncSetAttribute(paired_bf16_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[ctx->device & 63] = true;
  }
  const int64_t tiles = a.mtiles * a.ntiles;
  const int g = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
  const bool probed = mnf_probe_begin(ctx, MNF_PROBE_DENSE_GEMM);
  mnf_note_path("paired_bf16<128>");
  paired_bf16_fwd<<<g, THREADS, smem, ctx->stream>>>(a);
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}
