// Shared internals of libmnf_b200: context, error plumbing, Philox, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mnf_b200.h"

struct mnf_ctx {
  int device;
  cudaStream_t stream;
  bool owns_stream;
  int sm_count;
  int64_t launches;
  void* ws;  // grow-only device workspace
  size_t ws_bytes;
  bool capturing;
  int math_mode;  // MNF_MATH_FP32 (FFMA), MNF_MATH_TF32X3 (tcgen05, 3xTF32 split) or MNF_MATH_BF16 (dense: one bf16 MMA)
  // timing probe: bracket the probe_skip-th launch of kernel family probe_kind with events
  int probe_kind;
  int probe_skip;
  cudaEvent_t probe_start, probe_stop;
  // frozen weights (mnf_ctx_freeze_weights): packed / pre-split weight blocks are kept in
  // persistent device buffers keyed by the weight pointers instead of being rebuilt per call
  bool frozen;
  struct mnf_packed_entry* packed;
  int packed_n, packed_cap;
  // geometry tables (mnf_geom_table): step offsets / k maps that depend on a layer's SHAPE only; uploaded once
  // per context and kept whether or not the weights are frozen, so a captured training step holds no H2D copy
  struct mnf_packed_entry* geom;
  int geom_n, geom_cap;
};
struct mnf_packed_entry {
  uint64_t key[4];
  void* ptr;
  size_t bytes;
};

enum { MNF_PROBE_NONE = 0, MNF_PROBE_CONV_GEMM = 1, MNF_PROBE_DENSE_GEMM = 2, MNF_PROBE_FLOW_STEP = 3, MNF_PROBE_KL = 4,
       MNF_PROBE_MC_EXPAND = 5 };

// returns true when this launch is the probed one (and records the start event)
static inline bool mnf_probe_begin(mnf_ctx* ctx, int kind) {
  if (ctx->probe_kind != kind) return false;
  if (ctx->probe_skip-- != 0) return false;
  cudaEventRecord(ctx->probe_start, ctx->stream);
  return true;
}
static inline void mnf_probe_end(mnf_ctx* ctx, bool on) {
  if (on) {
    cudaEventRecord(ctx->probe_stop, ctx->stream);
    ctx->probe_kind = MNF_PROBE_NONE;
  }
}

void mnf_set_error(const char* fmt, ...);
struct NoiseDev;
int mnf_flow_umma_step(mnf_ctx* ctx, int64_t B, int D, int H, int kind, int parity, const float* zin,
                       const float* q0_mean, const float* q0_log_var, const NoiseDev& eps, float* z0_out, float* zout,
                       const float* ld_in, float* ld_out, float* hid_save, const mnf_flow_step& st,
                       const NoiseDev& bern, float* pack_ws);
size_t mnf_flow_umma_pack_floats(int D);
// flows_f16.cu: RNVP steps on kind::f16 (scaled fp16 hi/lo operands), z moved by bulk copies only
bool mnf_flow_f16_ok(const mnf_ctx* ctx, int64_t B, int64_t D, const mnf_flow_step& st);
size_t mnf_flow_f16_pack_bytes(int D);
size_t mnf_flow_f16_tiles(int64_t B);
int mnf_flow_f16_sample_z0(mnf_ctx* ctx, int64_t B, int D, const float* q0_mean, const float* q0_log_var, const NoiseDev& eps,
                           float* zout, float* tmax);
int mnf_flow_f16_absmax(mnf_ctx* ctx, int64_t B, int D, const float* z, float* tmax);
int mnf_flow_f16_step(mnf_ctx* ctx, int64_t B, int D, const float* zin, float* zout, const float* ld_in, float* ld_out,
                      float* hid_save, const mnf_flow_step& st, const NoiseDev& bern, const float* tmax_in, float* tmax_out,
                      uint8_t* pack_ws);
// flows_generic.cu: coupling networks with several hidden layers / other activations (warp per row, FFMA)
bool mnf_flow_step_is_generic(const mnf_flow_step& st);
int64_t mnf_flow_step_hidden_total(const mnf_flow_step& st);
int mnf_flow_generic_forward(mnf_ctx* ctx, int64_t B, int64_t D, const mnf_flow_step& st, const float* zin, float* zout,
                             const float* ld_in, float* ld_out, float* hid_save);
int mnf_flow_generic_backward(mnf_ctx* ctx, int64_t B, int64_t D, const mnf_flow_step& st, const mnf_flow_step& gs,
                              const float* xin, const float* hid, const float* dy, const float* dld, const float* add,
                              float* dx);
int mnf_conv_raw_umma_forward(mnf_ctx* ctx, int64_t B, const int* geo, int N, const float* x, const float* z,
                              const float* mean_W, const float* log_std_W, const float* mean_b,
                              const float* log_var_b, float max_std, const NoiseDev& eps, float* y, float* mean_out,
                              float* sd_out, bool* used);
// conv_raw_f16.cu: the raw-tile kernel on kind::f16 with scaled hi/lo fp16 operands
int mnf_conv_raw_f16_forward(mnf_ctx* ctx, int64_t B, const int* geo, int N, const float* x, const float* z,
                             const float* mean_W, const float* log_std_W, const float* mean_b,
                             const float* log_var_b, float max_std, const NoiseDev& eps, float* y, float* mean_out,
                             float* sd_out, bool* used, const mnf_mc_source* gen = nullptr);
// conv_win_umma.cu: single-channel stride-1 VALID convolutions (sliding-window planes)
int mnf_conv_win_umma_forward(mnf_ctx* ctx, int64_t B, const int* geo, int N, const float* x, const float* z,
                              const float* mean_W, const float* log_std_W, const float* mean_b,
                              const float* log_var_b, float max_std, const NoiseDev& eps, float* y, bool* used);
int mnf_tc_paired_forward(mnf_ctx* ctx, int conv, int64_t M, int K, int N, const float* x, const float* z,
                          const float* mean_W, const float* log_std_W, const float* mean_b, const float* log_var_b,
                          float max_std, const NoiseDev& eps, float* y, float* mean_out, float* sd_out,
                          const int* geo, int act = 0);  // conv geo: H,W,C,kh,kw,Ho,Wo,stride,pad_t,pad_l,pool; act 1 = ReLU (dense)
// dense_bf16.cu: MNF_MATH_BF16, one kind::f16 (bf16) MMA per product, TMA-fed from pre-packed operands
int mnf_dense_bf16_forward(mnf_ctx* ctx, int64_t M, int K, int N, const float* x, const float* z, const float* mean_W,
                           const float* log_std_W, const float* mean_b, const float* log_var_b, float max_std,
                           const NoiseDev& eps, float* y, float* sd_out, int act);
// dense_f16x3.cu: many-row dense layers in the fp32-parity mode: operands packed once into scaled fp16 hi/lo
// planes, three kind::f16 MMAs per product, both operand streams by cp.async.bulk
bool mnf_dense_f16x3_ok(const mnf_ctx* ctx, int64_t M, int K, int N);
int mnf_dense_f16x3_forward(mnf_ctx* ctx, int64_t M, int K, int N, const float* x, const float* z, const float* mean_W,
                            const float* log_std_W, const float* mean_b, const float* log_var_b, float max_std,
                            const NoiseDev& eps, float* y, float* sd_out, int act);
// gemm_umma.cu: strided fp32 GEMM C (+)= A B as 3xTF32 on tcgen05 (backward contractions)
int mnf_gemm_umma(mnf_ctx* ctx, int M, int N, int K, const float* A, int64_t sAm, int64_t sAk, const float* B,
                  int64_t sBk, int64_t sBn, float* C, int64_t ldc, bool accumulate, bool sqA = false);
// backward.cu: C (+)= A B with arbitrary strides (FFMA, or tcgen05 3xTF32 in the tensor-core math modes)
int mnf_gemm_any(mnf_ctx* ctx, int M, int N, int K, const float* A, int64_t sAm, int64_t sAk, const float* B,
                 int64_t sBk, int64_t sBn, float* C, int64_t ldc, bool accumulate, bool sqA = false);
void* mnf_workspace(mnf_ctx* ctx, size_t bytes);  // nullptr on failure (error set)
void mnf_note_path(const char* name);              // which kernel family served the last forward call (mnf_last_path)
// Where a call's packed weights go.  Not frozen: the transient workspace, *fresh = true (pack every
// call).  Frozen: a persistent block per key; *fresh tells whether it still has to be filled.
void* mnf_packed(mnf_ctx* ctx, uint64_t k0, uint64_t k1, uint64_t k2, uint64_t k3, size_t bytes, bool* fresh);
// device copy of n host ints, cached by (key, a checksum of the data, n); nullptr on failure (error set)
int* mnf_geom_table(mnf_ctx* ctx, uint64_t key, const int* h_data, size_t n);
static inline uint32_t __float_as_uint_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline uint64_t mnf_mix(uint64_t a, uint64_t b) { return (a ^ (b + 0x9E3779B97F4A7C15ull + (a << 6) + (a >> 2))); }

#define MNF_CHECK_ARG(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      mnf_set_error(__VA_ARGS__);     \
      return MNF_ERR_INVALID;         \
    }                                 \
  } while (0)

#define MNF_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      mnf_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return MNF_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define MNF_LAUNCHED(ctx)                                                           \
  do {                                                                              \
    (ctx)->launches++;                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      mnf_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return MNF_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// The tensor core adds each MMA's result to the fp32 accumulator in TMEM without round-to-nearest,
// so a chain of K/8 x 3 accumulations drifts: against the float64 oracle the 3xTF32 contraction is
// at 6e-6 of the tensor's maximum for K = 800 and 2.7e-5 for K = 4096 (FFMA: 9e-7 / 2.7e-6;
// tests/exp_accumulation_drift.py).  The fp32-parity mode therefore keeps contractions longer than this on
// the FFMA kernels (every MNF-LeNet / MLP layer is shorter); the bf16 mode, whose caller accepted a
// tolerance, uses the tensor cores for any K.
#define MNF_TC_PARITY_KMAX 1280
static inline bool mnf_tc_k_ok(const mnf_ctx* ctx, int64_t K) {
  return ctx->math_mode == MNF_MATH_BF16 || (ctx->math_mode == MNF_MATH_TF32X3 && K <= MNF_TC_PARITY_KMAX);
}

// ------------------------------------------------------------------------------- Philox
// Philox4x32-10 (Salmon et al. 2011).  counter = (c0,c1,c2,c3), key = (k0,k1).
struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                            uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// Device-side view of mnf_noise.
struct NoiseDev {
  uint32_t k0, k1;
  uint32_t stream;
  uint64_t row_offset;
  const float* injected;
  const uint32_t* delta;  // device counter added to k1 (the call index) at run time, or nullptr
};

static inline NoiseDev make_noise(const mnf_noise* n) {
  NoiseDev d;
  if (n) {
    d.k0 = (uint32_t)n->seed; d.k1 = (uint32_t)(n->seed >> 32);
    d.stream = n->stream; d.row_offset = n->row_offset; d.injected = n->injected; d.delta = n->call_delta;
  } else {
    d.k0 = d.k1 = 0; d.stream = 0; d.row_offset = 0; d.injected = nullptr; d.delta = nullptr;
  }
  return d;
}

#ifdef __CUDACC__
// counter layout documented in include/mnf_b200.h; rows beyond 2^32 spill into the top
// 8 bits of the stream word.
__device__ __forceinline__ Philox4 philox_at(const NoiseDev& n, uint64_t grow, uint32_t outer, uint32_t inner4) {
  uint32_t c3 = n.stream ^ ((uint32_t)(grow >> 32) << 24);
  const uint32_t k1 = n.delta ? n.k1 + __ldg(n.delta) : n.k1;  // captured graphs: fresh call index per replay
  return philox4x32_10(inner4, outer, (uint32_t)grow, c3, n.k0, k1);
}

__device__ __forceinline__ float u01_open(uint32_t x) {  // (0,1), 24 bits
  return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// MUFU.SQRT (2^-23 relative error, no denormal slow path); arguments here are >= 0 and normal
__device__ __forceinline__ float sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// MUFU.LG2 without the denormal pre-scaling of __log2f (three extra instructions per call): for normal
// arguments the result is the same bit pattern
__device__ __forceinline__ float lg2_fast(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Box-Muller on two 32-bit words -> two N(0,1).  Every noise consumer is issue-bound on this transform
// (conv_mc_expand: 118 M normals per step), so it is built from the cheapest instructions that keep the
// distribution exact to fp32:
//   word -> float without I2F (quarter-rate XU pipe, shared with MUFU): the top 23 bits become the
//           mantissa of f in [1, 2);  u = f - (1 - 2^-24) is (k + 1/2) 2^-23 in (0, 1), symmetric about 1/2;
//           the angle is 2 pi (f - 1.5), 2^23 equidistant angles in [-pi, pi);
//   radius  sqrt(-2 ln u) = sqrt(lg2(u) * (-2 ln 2)): MUFU.LG2, one FMUL, MUFU.SQRT;
//   sin / cos of the same argument share one range-reduction multiply (MUFU.SIN, MUFU.COS).
// Four MUFU, five FMUL, two FADD and four integer operations per pair (before: four MUFU + two I2F, ten
// FMUL, four FADD).  tests/gpu_util.py:box_muller_f64 restates it in float64.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float fa = __uint_as_float((a >> 9) | 0x3F800000u), fb = __uint_as_float((b >> 9) | 0x3F800000u);
  const float u = fa - 0.99999994f;  // 1 - 2^-24, exact in fp32
  const float r = sqrt_fast(lg2_fast(u) * -1.3862943611198906f);
  float s, c;
  __sincosf(6.283185307179586f * (fb - 1.5f), &s, &c);
  n0 = r * c;
  n1 = r * s;
}

__device__ __forceinline__ void philox_normal4(const NoiseDev& n, uint64_t grow, uint32_t outer, uint32_t inner4,
                                               float out[4]) {
  Philox4 p = philox_at(n, grow, outer, inner4);
  box_muller(p.x, p.y, out[0], out[1]);
  box_muller(p.z, p.w, out[2], out[3]);
}

// Round keys hoisted out of the call: the key schedule (k += W per round, 20 integer adds) does not
// depend on the counter, but with the call index read from memory (NoiseDev::delta) the compiler
// re-derives it inside every call.  A kernel that draws many blocks per thread computes the ten
// round keys once and keeps them in registers.
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
};
__device__ __forceinline__ PhiloxKeys philox_keys(const NoiseDev& n) {
  PhiloxKeys K;
  const uint32_t k1 = n.delta ? n.k1 + __ldg(n.delta) : n.k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    K.k0[r] = n.k0 + (uint32_t)r * 0x9E3779B9u;
    K.k1[r] = k1 + (uint32_t)r * 0xBB67AE85u;
  }
  return K;
}
__device__ __forceinline__ void philox_normal4_keys(const NoiseDev& n, const PhiloxKeys& K, uint64_t grow, uint32_t outer,
                                                    uint32_t inner4, float out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t c0 = inner4, c1 = outer, c2 = (uint32_t)grow, c3 = n.stream ^ ((uint32_t)(grow >> 32) << 24);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ K.k0[r], n2 = hi0 ^ c3 ^ K.k1[r];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  box_muller(c0, c1, out[0], out[1]);
  box_muller(c2, c3, out[2], out[3]);
}

// one normal at (row, outer, inner); `stride_*` describe the injected layout [rows, outer, inner]
__device__ __forceinline__ float noise_normal1(const NoiseDev& n, int64_t lrow, uint32_t outer, uint32_t inner,
                                               int64_t n_outer, int64_t n_inner) {
  if (n.injected) return n.injected[(lrow * n_outer + outer) * n_inner + inner];
  float v[4];
  philox_normal4(n, n.row_offset + (uint64_t)lrow, outer, inner >> 2, v);
  return v[inner & 3];
}

// Fair coins (p == 0.5, the RNVP mask) cost one BIT each: element d of a row is bit (d & 31) of
// word ((d >> 5) & 3) of the Philox block with inner counter d >> 7, and the draw is 1.0 when
// the bit is clear -- 128 draws per Philox call.  Any other p spends one word per element
// (u01(word) <= p, inner counter d >> 2).  include/mnf_b200.h documents both.
__device__ __forceinline__ uint32_t philox_word(const Philox4& q, uint32_t i) {
  return i == 0 ? q.x : i == 1 ? q.y : i == 2 ? q.z : q.w;
}
__device__ __forceinline__ float coin_of(const Philox4& q, uint32_t d) {
  return ((philox_word(q, (d >> 5) & 3u) >> (d & 31u)) & 1u) ? 0.0f : 1.0f;
}
__device__ __forceinline__ float noise_bern1(const NoiseDev& n, int64_t lrow, uint32_t inner, int64_t n_inner,
                                             float p) {
  if (n.injected) return n.injected[lrow * n_inner + inner];
  if (p == 0.5f) return coin_of(philox_at(n, n.row_offset + (uint64_t)lrow, 0u, inner >> 7), inner);
  Philox4 q = philox_at(n, n.row_offset + (uint64_t)lrow, 0u, inner >> 2);
  return u01_open(philox_word(q, inner & 3u)) <= p ? 1.0f : 0.0f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif
