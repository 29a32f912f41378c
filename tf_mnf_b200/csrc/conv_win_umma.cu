// MNFConv2D.call (mnf_conv.py:182-197) for SINGLE-CHANNEL inputs on tcgen05 (MNF-LeNet conv1:
// 28x28x1 -> 24x24x20, 5x5), again without materialising im2col.
//
// With one input channel the K axis of the contraction is the filter window itself.  A
// K-major SWIZZLE_NONE operand wants the 8 rows of a core matrix 16 bytes apart, so the planes
// hold *sliding windows*: chunk (hi, img, w) = x[img][hi][w .. w+3] (16 bytes, zero past the
// row end).  For output row r = ((ho, img), wo) and filter tap row i, the four taps j0..j0+3 are
// the chunk at (ho+i, img, wo+j0): the 8 pixels wo = 0..7 of a core matrix are 8 consecutive
// chunks, consecutive core matrices (next image, then next output row) are one plane row apart
// (SBO), and the two K chunks of one K=8 MMA (taps j = 0..3 and 4..7 of one filter row, or two
// filter rows when kw <= 4) are `lbo` bytes apart, taken per MMA from a host-built step table.
// Taps past the filter width carry zero weights.
//
//   planes[buf 2][v 4][hi H][img 2][w Wp] chunks, v in {x hi, x lo, x^2 hi, x^2 lo} (3xTF32);
//   one plane set = two images, double buffered: the next pair is split while the tensor
//   core works through the (Ho/8)*(Wo/8) tiles of this one.
//   tile (hb, wg): rows r = ((hl*2 + img)*8 + wl), ho = 8*hb + hl, wo = 8*wg + wl; M = 128.
//   weights: all steps of both phases stay resident in shared memory (N <= 32: 2 KB per step).
//   x_hi * [W_hi | W_lo] is one N=64 instruction, x_lo * W_hi one N=32 instruction; four TMEM
//   accumulator buffers of 128 columns ([mean 64 | var 64]) let the MMA warp run ahead.
//
//   warps 0-15   epilogue (the Philox + Box-Muller noise of 128x20 outputs per tile is the long
//                pole): four groups of four warps; group g owns TMEM accumulator buffer g and
//                the tiles t = g (mod 4), warp w of it reads lanes 32*(w%4).. and ALL column
//                quads of its rows (the per-tile preamble is paid once per tile, not once per
//                quad) -> bias, z-scale, eps, optional ReLU + 2x2 max-pool by warp shuffles (the
//                four pixels of a pool window sit in one warp), store
//   warp  16     MMA issue;   warps 17-19: plane producers (one lane per chunk, the three
//                neighbours of a window by shuffle)
#include <vector>

#include "umma.cuh"

namespace cw {
using namespace um;

constexpr int BM = 128, BN = 32;
constexpr int STEP_B = 2 * 2 * BN * 16;  // bytes of one step's weights for one phase: [2 kc][hi 32 n | lo 32 n][16 B]
constexpr int EPI_E = 4;                     // epilogue groups = TMEM accumulator buffers
constexpr int EPI_WARPS = 4 * EPI_E, MMA_WARP = EPI_WARPS, THREADS = (EPI_WARPS + 4) * 32;
constexpr int PROD_WARPS = 3;
constexpr int NACC = 4, ACC_COLS = 4 * BN;  // [mean: hi-product 32 | lo-product 32][var: same]
constexpr int TMEM_COLS = NACC * ACC_COLS;  // 512
constexpr int I = 2;                         // images per plane set
constexpr int MAXS = 8;                      // K=8 steps per phase (filter window <= 64 taps incl. padding)
constexpr int RB = 7;                        // plane rows a producer warp keeps in flight

struct Args {
  int64_t B;
  int H, W, Ho, Wo, N, Wp;  // Wp: chunks per plane row (<= 32)
  int steps, pool;
  const float* x;
  const float* z;  // [B,N] or nullptr
  const float* bm;
  const float* vb;
  const int* steptab;  // [steps][2]: a_off, a_lbo (bytes)
  const float* Bp;     // [2 phases][steps][STEP_B/4 floats]
  float* y;
  NoiseDev eps;
  uint64_t dA[8];  // A descriptor of every step without the tile base (constant bank -> uniform registers)
};

// the destination registers are written asynchronously: nothing may touch them before tmem_ld_wait()
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float v[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
               : "r"(taddr));
}
// tcgen05.mma without the "memory" clobber: volatile keeps it ordered against the mbarrier waits
// and commits (also volatile), while ordinary address arithmetic may move across it
__device__ __forceinline__ void umma_tf32_nc(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// INJ: the output noise is read from memory (tests); compiled apart so that the epilogue's Philox path is branch-free
template <bool INJ>
__global__ void __launch_bounds__(THREADS, 1) conv_win_umma(const __grid_constant__ Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  const int PV = a.H * I * a.Wp * 16;  // bytes of one plane variant
  uint8_t* planes = smem;                        // [2][4][PV]
  uint8_t* Bs = planes + 8 * PV;                 // [2 phases][steps][STEP_B]
  int* stab = (int*)(Bs + 2 * a.steps * STEP_B);  // [steps][2]
  float* bias_s = (float*)(((uintptr_t)(stab + 2 * a.steps) + 15) & ~(uintptr_t)15);  // [2][BN]
  uint64_t* bars = (uint64_t*)(((uintptr_t)(bias_s + 2 * BN) + 7) & ~(uintptr_t)7);
  uint64_t* pfull = bars;        // [2]
  uint64_t* pempty = bars + 2;   // [2]
  uint64_t* tfull = bars + 4;    // [NACC]
  uint64_t* tempty = bars + 8;   // [NACC]
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);

  const int nhb = a.Ho >> 3, nwg = a.Wo >> 3, ntile = nhb * nwg;
  const int64_t nsets = (a.B + I - 1) / I;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(&pfull[b], PROD_WARPS); mbar_init(&pempty[b], 1); }
    for (int b = 0; b < NACC; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // resident weights, step table, biases
  {
    const float4* src = reinterpret_cast<const float4*>(a.Bp);
    float4* dst = reinterpret_cast<float4*>(Bs);
    const int n16 = 2 * a.steps * STEP_B / 16;
    for (int i = threadIdx.x; i < n16; i += THREADS) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < 2 * a.steps; i += THREADS) stab[i] = a.steptab[i];
    for (int n = threadIdx.x; n < BN; n += THREADS) {
      bias_s[n] = n < a.N ? a.bm[n] : 0.f;
      bias_s[BN + n] = n < a.N ? a.vb[n] : 1.f;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp < EPI_WARPS) {
    // ================================================================== epilogue warps
    // The kernel is issue-bound on this code (Philox4x32-10 + Box-Muller per 4 outputs).
    const int quad = warp & 3, grp = warp >> 2;    // grp == accumulator buffer
    const int nq = a.N >> 2;                       // column quads (N % 4 == 0)
    const int gl = lane >> 3, img = gl & 1, hl = quad * 2 + (gl >> 1), wl = lane & 7;
    const bool writer = !a.pool || (lane & 17) == 0;
    const int HoWo = a.Ho * a.Wo;
    // element offsets of this lane's pixel inside one image's output, for tile (0,0), and the
    // strides of the tile grid
    const int outer0 = hl * a.Wo + wl;
    const int pix0 = a.pool ? (hl >> 1) * (a.Wo >> 1) + (wl >> 1) : outer0;
    const int pix_wg = a.pool ? 4 : 8, pix_hb = a.pool ? 4 * (a.Wo >> 1) : 8 * a.Wo;
    const int64_t img_out = a.pool ? (int64_t)(a.Ho >> 1) * (a.Wo >> 1) * a.N : (int64_t)HoWo * a.N;
    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(grp * ACC_COLS);
    const int64_t my_sets = (nsets > blockIdx.x) ? (nsets - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_tiles = my_sets * ntile;
    uint32_t k = 0;  // this group's tile counter (parity of its accumulator barrier)
    const PhiloxKeys EK = philox_keys(a.eps);  // round keys once per warp (warp-uniform: uniform registers)
    for (int64_t tc = grp; tc < total_tiles; tc += EPI_E, ++k) {
      const int64_t si = tc / ntile;
      const int tile = (int)(tc - si * ntile);
      const int hb = tile / nwg, wg = tile - hb * nwg;
      const int64_t braw = (blockIdx.x + si * gridDim.x) * I + img;
      const bool ok = braw < a.B;
      const int64_t brow = ok ? braw : a.B - 1;
      const uint64_t grow = a.eps.row_offset + (uint64_t)brow;
      const bool st = ok && writer;
      const int outer = outer0 + hb * 8 * a.Wo + wg * 8, pix = pix0 + hb * pix_hb + wg * pix_wg;
      float* const dst = a.y + brow * img_out + (int64_t)pix * a.N;
      const float* const inj = a.eps.injected ? a.eps.injected + (brow * HoWo + outer) * a.N : nullptr;
      const float* const zrow = a.z ? a.z + brow * a.N : nullptr;
      mbar_wait(&tfull[grp], k & 1);
      tc_fence_after();
      float mh[4], ml[4], vh[4], vl[4];
      tmem_ld4(trow, mh);
      tmem_ld4(trow + BN, ml);
      tmem_ld4(trow + 2 * BN, vh);
      tmem_ld4(trow + 3 * BN, vl);
#pragma unroll 1
      for (int q = 0; q < nq; ++q) {
        const int nb = 4 * q;
        tmem_ld_wait();
        float mv[4], vv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mv[j] = mh[j] + ml[j];
          vv[j] = vh[j] + vl[j];
        }
        if (q + 1 < nq) {  // the next quad's accumulators travel while this one is computed
          tmem_ld4(trow + nb + 4, mh);
          tmem_ld4(trow + nb + 4 + BN, ml);
          tmem_ld4(trow + nb + 4 + 2 * BN, vh);
          tmem_ld4(trow + nb + 4 + 3 * BN, vl);
        } else {           // everything is in registers: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[grp]);
        }
        float e[4];
        if (INJ) {
          const float4 t = *reinterpret_cast<const float4*>(inj + nb);
          e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w;
        } else {
          philox_normal4_keys(a.eps, EK, grow, (uint32_t)outer, (uint32_t)q, e);
        }
        float4 zq = make_float4(1.f, 1.f, 1.f, 1.f);
        if (zrow) zq = __ldg(reinterpret_cast<const float4*>(zrow + nb));
        const float zz[4] = {zq.x, zq.y, zq.z, zq.w};
        const float4 bmq = *reinterpret_cast<const float4*>(bias_s + nb);
        const float4 bvq = *reinterpret_cast<const float4*>(bias_s + BN + nb);
        const float bm4[4] = {bmq.x, bmq.y, bmq.z, bmq.w}, bv4[4] = {bvq.x, bvq.y, bvq.z, bvq.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float mean = (mv[j] + bm4[j]) * zz[j];
          const float sd = sqrt_fast(vv[j] + bv4[j]);
          o[j] = fmaf(sd, e[j], mean);
        }
        if (a.pool) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v = fmaxf(o[j], __shfl_xor_sync(0xffffffffu, o[j], 1));   // wo pair
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));              // ho pair
            o[j] = fmaxf(v, 0.f);
          }
        }
        if (st) *reinterpret_cast<float4*>(dst + nb) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  } else if (warp == MMA_WARP) {
    // ======================================================================= MMA issuer
    // This warp shares its scheduler with five busy epilogue warps, so the issue loop is kept to
    // a handful of instructions per step: the per-step parts of the descriptors (offset, LBO,
    // SBO, version) come pre-built from the constant bank, the warp stays converged (elect inside
    // the issue wrappers, umma.cuh) and per MMA only the tile base is added -- in uniform registers.
    const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint64_t dB0 = make_desc(smem_u32(Bs), 2 * BN * 16, 128);
    const uint32_t plane0 = smem_u32(planes) >> 4, pv16 = (uint32_t)PV >> 4;
    const uint32_t bph = (uint32_t)(a.steps * STEP_B) >> 4;  // phase 1 weights follow phase 0
    uint32_t tcount = 0;
    int it = 0;
    for (int64_t s = blockIdx.x; s < nsets; s += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&pfull[buf], (it >> 1) & 1);
      tc_fence_after();
      int hb = 0, wg = 0;
      for (int tile = 0; tile < ntile; ++tile, ++tcount) {
        const int acc = tcount & (NACC - 1);
        mbar_wait(&tempty[acc], ((tcount / NACC) & 1) ^ 1);
        tc_fence_after();
        {
          const uint32_t tbase = plane0 + (uint32_t)(buf * 4) * pv16 + (uint32_t)(hb * 8 * I * a.Wp + wg * 8);
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            const uint32_t dacc = tmem_base + (uint32_t)(acc * ACC_COLS + ph * 2 * BN);
            const uint64_t abase = (uint64_t)(tbase + (uint32_t)(2 * ph) * pv16);
#pragma unroll
            for (int k = 0; k < MAXS; ++k) {
              if (k < a.steps) {
                const uint64_t ah = a.dA[k] + abase, al = ah + pv16;
                const uint64_t bh = dB0 + (uint64_t)((k * STEP_B) >> 4) + (ph ? bph : 0u);
                umma_tf32_elect(dacc, ah, bh, idesc2, k == 0 ? 0u : 1u);  // x_hi * [W_hi | W_lo]
                umma_tf32_elect(dacc, al, bh, idesc1, 1u);                // x_lo * W_hi
              }
            }
          }
          umma_commit_elect(&tfull[acc]);
          if (tile == ntile - 1) umma_commit_elect(&pempty[buf]);  // this pair's planes may be refilled
        }
        __syncwarp();
        if (++wg == nwg) { wg = 0; ++hb; }
      }
    }
  } else {
    // ================================================================== plane producers
    const int pw = warp - MMA_WARP - 1;  // 0..2
    const int nrows = a.H * I;                 // plane rows (hi, img), img fastest
    int it = 0;
    for (int64_t s = blockIdx.x; s < nsets; s += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&pempty[buf], ((it >> 1) & 1) ^ 1);
      uint8_t* pb = planes + (size_t)buf * 4 * PV;
      for (int r0 = pw; r0 < nrows; r0 += PROD_WARPS * RB) {
        float v[RB];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int r = r0 + u * PROD_WARPS;
          v[u] = 0.f;
          if (r < nrows && lane < a.W) {
            int64_t b = s * I + (r & 1);
            if (b >= a.B) b = a.B - 1;
            v[u] = __ldg(a.x + (b * a.H + (r >> 1)) * a.W + lane);
          }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int r = r0 + u * PROD_WARPS;
          if (r >= nrows) break;  // warp-uniform
          float p[4][4];          // [variant][window element]
          split_tf32(v[u], p[0][0], p[1][0]);
          split_tf32(v[u] * v[u], p[2][0], p[3][0]);
#pragma unroll
          for (int vi = 0; vi < 4; ++vi)
#pragma unroll
            for (int d = 1; d < 4; ++d) {
              const float t = __shfl_down_sync(0xffffffffu, p[vi][0], d);
              p[vi][d] = lane + d < 32 ? t : 0.f;  // lanes >= W hold zeros already
            }
          if (lane < a.Wp) {
            uint8_t* dst = pb + (r * a.Wp + lane) * 16;
#pragma unroll
            for (int vi = 0; vi < 4; ++vi)
              *reinterpret_cast<float4*>(dst + vi * PV) = make_float4(p[vi][0], p[vi][1], p[vi][2], p[vi][3]);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// weights -> Bp[phase][step][kc 2][hi n 32 | lo n 32][4 k]; kmap[step][8] = filter element or -1
__global__ void pack_win_weights(int N, int steps, const int* __restrict__ kmap, const float* __restrict__ Wm,
                                 const float* __restrict__ Ls, float max_std, float* __restrict__ Bp,
                                 const float* __restrict__ lvb, float* __restrict__ vb) {
  const int total = 2 * steps * 2 * BN * 4;  // (phase, step, kc, n, e)
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int e = idx & 3;
    int t = idx >> 2;
    const int n = t % BN;
    t /= BN;
    const int kc = t & 1;
    t >>= 1;
    const int s = t % steps, ph = t / steps;
    const int k = kmap[s * 8 + kc * 4 + e];
    float w = 0.f;
    if (k >= 0 && n < N) {
      if (ph == 0) w = Wm[k * N + n];
      else {
        const float sd = fminf(expf(Ls[k * N + n]), max_std);
        w = sd * sd;
      }
    }
    float hi, lo;
    split_tf32(w, hi, lo);
    float* base = Bp + (ph * steps + s) * (STEP_B / 4) + (kc * 2 * BN + n) * 4 + e;
    base[0] = hi;
    base[BN * 4] = lo;
    if (idx < N) vb[idx] = fminf(expf(lvb[idx]), max_std * max_std);
  }
}

}  // namespace cw

// Returns MNF_OK with *used = false when the shape is outside this kernel's domain.
int mnf_conv_win_umma_forward(mnf_ctx* ctx, int64_t B, const int* geo, int N, const float* x, const float* z,
                              const float* mean_W, const float* log_std_W, const float* mean_b,
                              const float* log_var_b, float max_std, const NoiseDev& eps, float* y, bool* used) {
  using namespace cw;
  *used = false;
  const int H = geo[0], W = geo[1], C = geo[2], kh = geo[3], kw = geo[4], Ho = geo[5], Wo = geo[6], stride = geo[7],
            pad_t = geo[8], pad_l = geo[9], pool = geo[10];
  if (C != 1 || stride != 1 || pad_t || pad_l || (Ho & 7) || (Wo & 7) || (N & 3) || N > BN || N < 4) return MNF_OK;
  if (Ho != H - kh + 1 || Wo != W - kw + 1 || W > 32) return MNF_OK;  // VALID geometry, one lane per column
  if (z && ((uintptr_t)z & 15)) return MNF_OK;
  if (((uintptr_t)y & 15) || (eps.injected && ((uintptr_t)eps.injected & 15))) return MNF_OK;
  Args a = {};
  a.B = B; a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo; a.N = N; a.pool = pool;
  a.Wp = Wo + 4 * ((kw - 1) / 4);
  if (a.Wp > 32) return MNF_OK;
  // ---- step table: the 4-tap chunks of the filter window in plane order, two per K=8 step
  std::vector<int> off, tab, kmap;
  std::vector<int> ck;  // first filter element of each chunk (row i, tap j0), -1 = padding chunk
  for (int i = 0; i < kh; ++i)
    for (int j0 = 0; j0 < kw; j0 += 4) {
      off.push_back((i * I * a.Wp + j0) * 16);
      ck.push_back(i * kw + j0);
    }
  auto push_k = [&](int c) {
    for (int e = 0; e < 4; ++e) {
      const bool valid = c >= 0 && (ck[c] % kw) + e < kw;
      kmap.push_back(valid ? ck[c] + e : -1);
    }
  };
  for (size_t c = 0; c < off.size(); c += 2) {
    tab.push_back(off[c]);
    push_k((int)c);
    if (c + 1 < off.size()) {
      tab.push_back(off[c + 1] - off[c]);
      push_k((int)c + 1);
    } else {
      tab.push_back(16);  // dummy second chunk: valid memory, zero weights
      push_k(-1);
    }
  }
  a.steps = (int)tab.size() / 2;
  if (a.steps > MAXS) return MNF_OK;
  for (int i = 0; i < a.steps; ++i)  // K-major SWIZZLE_NONE descriptor (umma.cuh:make_desc) with start = step offset
    a.dA[i] = (uint64_t)(((uint32_t)tab[2 * i] >> 4) & 0x3FFF) | ((uint64_t)(((uint32_t)tab[2 * i + 1] >> 4) & 0x3FFF) << 16) |
              ((uint64_t)(((uint32_t)(a.Wp * 16) >> 4) & 0x3FFF) << 32) | (1ull << 46);
  const size_t PV = (size_t)H * I * a.Wp * 16;
  const size_t smem = 8 * PV + (size_t)2 * a.steps * STEP_B + sizeof(int) * 2 * a.steps + sizeof(float) * 2 * BN + 256;
  if (smem > 227 * 1024) return MNF_OK;
  // ---- workspace: Bp | vb | steptab | kmap
  const size_t bp_floats = (size_t)2 * a.steps * (STEP_B / 4);
  const size_t ws_bytes = sizeof(float) * (bp_floats + N + 64) + sizeof(int) * (tab.size() + kmap.size() + 128);
  bool fresh = true;
  char* ws = (char*)mnf_packed(ctx, mnf_mix(0xC017ull, mnf_mix(((uint64_t)H << 32) | (uint32_t)W, ((uint64_t)kh << 40) | ((uint64_t)kw << 20) | (uint32_t)N)),
                               (uint64_t)(uintptr_t)mean_W, (uint64_t)(uintptr_t)log_std_W,
                               mnf_mix((uint64_t)(uintptr_t)log_var_b, (uint64_t)__float_as_uint_host(max_std)), ws_bytes, &fresh);
  if (!ws) return MNF_ERR_CUDA;
  float* Bp = (float*)ws;
  float* vb = Bp + bp_floats;
  int* d_tab = mnf_geom_table(ctx, 0xC017A, tab.data(), tab.size());  // shape-only tables: uploaded once per context
  int* d_kmap = mnf_geom_table(ctx, 0xC017B, kmap.data(), kmap.size());
  if (!d_tab || !d_kmap) return MNF_ERR_CUDA;
  if (fresh) {
    const int tot = 2 * a.steps * 2 * BN * 4;
    pack_win_weights<<<cdiv(tot, 256), 256, 0, ctx->stream>>>(N, a.steps, d_kmap, mean_W, log_std_W, max_std, Bp, log_var_b, vb);
    MNF_LAUNCHED(ctx);
  }
  a.x = x; a.z = z; a.bm = mean_b; a.vb = vb; a.steptab = d_tab; a.Bp = Bp; a.y = y; a.eps = eps;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MNF_CUDA(cudaFuncSetAttribute(conv_win_umma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MNF_CUDA(cudaFuncSetAttribute(conv_win_umma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const int64_t sets = cdiv(B, (int64_t)I);
  const int g = (int)(sets < ctx->sm_count ? sets : ctx->sm_count);
  const bool probed = mnf_probe_begin(ctx, MNF_PROBE_CONV_GEMM);
  mnf_note_path("conv_win_umma");
  if (eps.injected) conv_win_umma<true><<<g, THREADS, smem, ctx->stream>>>(a);
  else conv_win_umma<false><<<g, THREADS, smem, ctx->stream>>>(a);
  mnf_probe_end(ctx, probed);
  MNF_LAUNCHED(ctx);
  *used = true;
  return MNF_OK;
}
