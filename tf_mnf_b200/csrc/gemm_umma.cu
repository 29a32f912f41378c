// Strided fp32 GEMM on tcgen05 for the BACKWARD contractions (dA = dy W^T,
// dW = (x*z)^T dy, the variance twins, and the im2col'ed conv versions): C (+)= A B with
// A(m,k) = A[m sAm + k sAk], B(k,n) = B[k sBk + n sBn], any strides, so that transposed operands
// need no copy.  fp32 parity as in tc_gemm.cu: both operands are split hi + lo (round-to-nearest
// TF32) and each product is hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM.
//
// One 128 x 128 tile (x one K slice) per CTA, two CTAs resident per SM (96 KB smem, 128 TMEM
// columns each), each with a three-deep register prefetch ring in its producers:
//   warps 0-7 producers: 8 + 8 operand elements per thread and 16-wide k-block, loaded three
//             blocks ahead, split, stored K-major (no-swizzle UMMA layout); thread -> element
//             mapping follows the operand's unit-stride axis so the loads coalesce either way;
//             afterwards the same warps are the epilogue (tcgen05.ld, store / atomic add)
//   warp 8    MMA issue (warp-converged, elect.sync): 6 x tcgen05.mma.kind::tf32 per k-block,
//             instruction N = the tile's real column count rounded up to 16
// Split-K over gridDim.z (atomic adds into a zeroed C) when there are fewer tiles than CTA slots.
#include "umma.cuh"

namespace gu {
using namespace um;

constexpr int BM = 128, BN = 128, BK = 16, KC = BK / 4;
constexpr int VAR_BYTES = KC * 128 * 16;    // hi or lo part of one operand tile: 8 KB
constexpr int OP_BYTES = 2 * VAR_BYTES;
constexpr int STAGE_BYTES = 2 * OP_BYTES;   // A hi, A lo, B hi, B lo: 32 KB
constexpr int STAGES = 3;                   // == depth of the producers' register prefetch ring
constexpr int PROD_WARPS = 8;
constexpr int THREADS = (PROD_WARPS + 1) * 32;
constexpr int TMEM_COLS = 128;

struct Args {
  int M, N, K;
  const float* A; int64_t sAm, sAk;
  const float* B; int64_t sBk, sBn;
  float* C; int64_t ldc;
  int atomic;   // accumulate into C (split-K or C += ...)
  int kchunk;   // K range per z-slice (multiple of BK)
  int sqA;      // multiply with A(m,k)^2 (variance twin of a contraction over the same operand)
};

// 8 elements of a [128 rows x 16 k] operand tile per thread.  kfast: the operand's k axis has unit
// stride -> 4 lanes cover the 16 k of a row; else rows are (or may be) contiguous -> a lane per row.
struct Frag {
  float v[8];
};
// FULLK: the whole 16-wide k-block lies inside the slice (every block but possibly the last): no
// per-element k predicate.  vec4: unit-stride k rows that are 16-byte aligned load as one float4.
// Addresses advance by pointer increments; the producers are instruction-bound otherwise.
template <bool FULLK>
__device__ __forceinline__ void load_frag(Frag& f, const float* __restrict__ base, int64_t s_row, int64_t s_k,
                                          int row0, int nrows, int rows_used, int k0, int kend, bool kfast, bool vec4,
                                          int pt) {
  if (kfast) {
    const int kc = pt & 3, r0 = pt >> 2, kk = k0 + kc * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rl = r0 + 64 * h, row = row0 + rl;
      const bool rok = rl < rows_used && row < nrows;
      const float* p = base + (int64_t)row * s_row + kk;  // s_k == 1
      if (rok && vec4 && (FULLK || kk + 4 <= kend)) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        f.v[h * 4 + 0] = t.x; f.v[h * 4 + 1] = t.y; f.v[h * 4 + 2] = t.z; f.v[h * 4 + 3] = t.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) f.v[h * 4 + e] = (rok && (FULLK || kk + e < kend)) ? __ldg(p + e) : 0.f;
      }
    }
  } else {
    const int rl = pt & 127, kh = pt >> 7, row = row0 + rl, kk = k0 + kh * 8;
    const bool rok = rl < rows_used && row < nrows;
    const float* p = base + (int64_t)row * s_row + (int64_t)kk * s_k;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      f.v[e] = (rok && (FULLK || kk + e < kend)) ? __ldg(p) : 0.f;
      p += s_k;
    }
  }
}
__device__ __forceinline__ void store_frag(const Frag& f, uint8_t* dst, bool kfast, int pt) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int kc, rl;
    if (kfast) { kc = pt & 3; rl = (pt >> 2) + 64 * h; } else { kc = 2 * (pt >> 7) + h; rl = pt & 127; }
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_tf32(f.v[h * 4 + e], hi[e], lo[e]);
    uint8_t* d = dst + (kc * 128 + rl) * 16;
    *reinterpret_cast<float4*>(d) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(d + VAR_BYTES) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void __launch_bounds__(THREADS, 2) gemm_tf32x3(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* done = empty + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(done + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * a.kchunk, kend = (kbeg + a.kchunk < a.K) ? kbeg + a.kchunk : a.K;
  const int nblk = (kend - kbeg + BK - 1) / BK;
  int n_eff = a.N - n0 < BN ? a.N - n0 : BN;
  n_eff = (n_eff + 15) & ~15;  // instruction N: multiple of 16

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], PROD_WARPS);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == PROD_WARPS) {
    // ======================================================================= MMA issuer
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_eff >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint64_t desc0 = make_desc(smem_u32(smem), 128 * 16, 128);
    for (int it = 0; it < nblk; ++it) {
      const int stage = it % STAGES;
      mbar_wait(&full[stage], (uint32_t)((it / STAGES) & 1));
      tc_fence_after();
      const uint64_t sd = desc0 + (uint64_t)((stage * STAGE_BYTES) >> 4);
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {  // one K = 8 instruction = two 16-byte chunks
        const uint64_t ah = sd + (uint64_t)((2 * j * 128 * 16) >> 4), al = ah + (VAR_BYTES >> 4);
        const uint64_t bh = ah + (OP_BYTES >> 4), bl = bh + (VAR_BYTES >> 4);
        umma_tf32_elect(tmem_base, ah, bh, idesc, (it == 0 && j == 0) ? 0u : 1u);
        umma_tf32_elect(tmem_base, al, bh, idesc, 1u);
        umma_tf32_elect(tmem_base, ah, bl, idesc, 1u);
      }
      umma_commit_elect(&empty[stage]);
      __syncwarp();
    }
    umma_commit_elect(done);
    __syncwarp();
  } else {
    // ======================================================================== producers
    const int pt = threadIdx.x;  // 0..255
    const bool a_kfast = a.sAk == 1, b_kfast = a.sBk == 1;
    // register ring: the loads of k-blocks it+1 .. it+STAGES-1 are in flight while block it is split and
    // stored (a synchronous load -> store loop pays one L2 round trip, ~2 us under load, per block)
    Frag fa[STAGES], fb[STAGES];
    const bool a_vec = a_kfast && (a.sAm & 3) == 0 && ((uintptr_t)a.A & 15) == 0;
    const bool b_vec = b_kfast && (a.sBn & 3) == 0 && ((uintptr_t)a.B & 15) == 0;
    auto issue = [&](Frag& xa, Frag& xb, int it) {
      const int k0 = kbeg + it * BK;
      if (k0 + BK <= kend) {
        load_frag<true>(xa, a.A, a.sAm, a.sAk, m0, a.M, BM, k0, kend, a_kfast, a_vec, pt);
        load_frag<true>(xb, a.B, a.sBn, a.sBk, n0, a.N, n_eff, k0, kend, b_kfast, b_vec, pt);
      } else {
        load_frag<false>(xa, a.A, a.sAm, a.sAk, m0, a.M, BM, k0, kend, a_kfast, a_vec, pt);
        load_frag<false>(xb, a.B, a.sBn, a.sBk, n0, a.N, n_eff, k0, kend, b_kfast, b_vec, pt);
      }
    };
#pragma unroll
    for (int d = 0; d < STAGES; ++d)
      if (d < nblk) issue(fa[d], fb[d], d);
    for (int it0 = 0; it0 < nblk; it0 += STAGES) {
#pragma unroll
      for (int d = 0; d < STAGES; ++d) {  // it0 is a multiple of STAGES: stage == ring slot == d
        const int it = it0 + d;
        if (it < nblk) {
          if (a.sqA) {
#pragma unroll
            for (int e = 0; e < 8; ++e) fa[d].v[e] *= fa[d].v[e];
          }
          mbar_wait(&empty[d], (uint32_t)(((it / STAGES) & 1) ^ 1));
          uint8_t* st = smem + d * STAGE_BYTES;
          store_frag(fa[d], st, a_kfast, pt);
          store_frag(fb[d], st + OP_BYTES, b_kfast, pt);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[d]);
          if (it + STAGES < nblk) issue(fa[d], fb[d], it + STAGES);
        }
      }
    }
    // ========================================================================= epilogue
    mbar_wait(done, 0u);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;  // TMEM lane quarter of this warp, column half
    const int m = m0 + q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool vec_ok = (a.ldc & 3) == 0 && ((uintptr_t)a.C & 15) == 0;
#pragma unroll 1
    for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 16) {
      if (c0 >= n_eff) break;  // warp-uniform
      float acc[16];
      tmem_ld16(trow + c0, acc);
      if (m < a.M) {
        float* crow = a.C + (int64_t)m * a.ldc + n0 + c0;
        if (!a.atomic && vec_ok && n0 + c0 + 16 <= a.N) {
          // 64 contiguous bytes of this thread's row: four 16-byte stores fill two whole sectors
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(crow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c0 + j < a.N) {
              if (a.atomic) atomicAdd(crow + j, acc[j]);
              else crow[j] = acc[j];
            }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == PROD_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

}  // namespace gu

// backward.cu's gemm() routes here when the context's math mode is a tensor-core one.
int mnf_gemm_umma(mnf_ctx* ctx, int M, int N, int K, const float* A, int64_t sAm, int64_t sAk, const float* B,
                  int64_t sBk, int64_t sBn, float* C, int64_t ldc, bool accumulate, bool sqA) {
  using namespace gu;
  if (M <= 0 || N <= 0) return MNF_OK;
  Args a = {M, N, K, A, sAm, sAk, B, sBk, sBn, C, ldc, accumulate ? 1 : 0, 0, sqA ? 1 : 0};
  dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(N, BN), 1);
  const int64_t tiles = (int64_t)grid.x * grid.y, slots = 2ll * ctx->sm_count;
  int split = 1;
  if (tiles < slots && K > 8 * BK && (accumulate || ldc == N)) {
    split = (int)(slots / tiles);
    const int maxsplit = (int)cdiv(K, 8 * BK);  // at least 8 k-blocks per slice
    if (split > maxsplit) split = maxsplit;
    if (split < 1) split = 1;
  }
  // the tensor core's fp32 accumulation is not round-to-nearest: keep every TMEM accumulation chain
  // to <= 1024 terms (error stays below 1e-5 of the maximum) and let fp32 atomics combine the slices
  if ((accumulate || ldc == N) && cdiv(K, split) > 1024) split = (int)cdiv(K, 1024);
  a.kchunk = (int)(cdiv(cdiv(K > 0 ? K : 1, split), BK) * BK);
  grid.z = (unsigned)cdiv(K > 0 ? K : 1, a.kchunk);
  if (grid.z > 1 && !accumulate) {
    MNF_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, ctx->stream));
    a.atomic = 1;
  }
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 128;
  static bool attr_set[64] = {};  // per device: function attributes belong to the device's copy of the kernel
  if (!attr_set[ctx->device & 63]) {
    MNF_CUDA(cudaFuncSetAttribute(gemm_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[ctx->device & 63] = true;
  }
  gemm_tf32x3<<<grid, THREADS, smem, ctx->stream>>>(a);
  MNF_LAUNCHED(ctx);
  return MNF_OK;
}
