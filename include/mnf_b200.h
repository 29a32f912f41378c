/*
 * mnf_b200.h -- C ABI of libmnf_b200.so: the MNF-layer hot path of janosh/tf-mnf as
 * hand-written CUDA for sm_100a (NVIDIA B200).
 *
 * The reference has no FFI of its own (it is pure Python on TensorFlow); the boundary it
 * exposes is the Keras-style class protocol.  Each entry point below therefore cites the
 * reference METHOD whose arithmetic it replaces (paths relative to the reference repo);
 * the Python host package tf_mnf_b200 binds these symbols with ctypes and re-creates the
 * tf_mnf.layers / tf_mnf.flows classes on top (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch/TF types.
 *   - every function returns 0 on success or a negative mnf_status; the message is
 *     available per thread from mnf_last_error().
 *   - all tensor pointers are DEVICE pointers to compact row-major float32 unless the
 *     parameter name starts with `h_` (host).  NULL is only legal where stated.
 *   - work is enqueued on the context's stream; nothing synchronises unless stated.
 *   - there is no CPU fallback: without a CUDA device mnf_ctx_create fails.
 *
 * Random numbers: counter-based Philox4x32-10.  A draw is addressed by
 *   key = seed (64 bit), counter = (inner >> 2, outer, global_row, stream), lane = inner & 3
 * so results do not depend on how rows are sharded over GPUs.  Every draw can instead be
 * INJECTED (mnf_noise.injected != NULL) which is how parity tests feed the oracle's noise.
 */
#ifndef MNF_B200_H
#define MNF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* only the functions declared here are exported from libmnf_b200.so */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define MNF_B200_VERSION 100 /* 0.1.0 */

typedef enum {
  MNF_OK = 0,
  MNF_ERR_INVALID = -1,     /* bad shape / null pointer / unsupported option -> ValueError */
  MNF_ERR_CUDA = -2,        /* CUDA runtime error -> RuntimeError */
  MNF_ERR_UNSUPPORTED = -3, /* valid in the reference, not implemented here (loud, never a fallback) */
  MNF_ERR_NOT_CUDA = -4     /* tensor is not on a CUDA device */
} mnf_status;

typedef struct mnf_ctx mnf_ctx; /* opaque: device, stream, workspace */

/* ---- runtime ------------------------------------------------------------------------ */
int mnf_version(void);
const char* mnf_last_error(void);
/* Name of the kernel family that served the calling thread's last dense / conv / flow forward
 * call ("conv_win_umma", "conv_raw_f16", "conv_raw_umma", "paired_umma", "conv_direct", "paired_gemm",
 * "flow_umma", "flow_mlp", ...): lets tests assert which path a shape exercised. */
const char* mnf_last_path(void);
int mnf_device_count(int* out);
/* stream: a cudaStream_t owned by the caller, or NULL to let the context create its own.
 * One device per process (one process per GPU): a context for a second device in the same
 * process is refused with MNF_ERR_UNSUPPORTED. */
int mnf_ctx_create(int device, void* stream, mnf_ctx** out);
/* A sibling context on the parent's device (own stream, own workspace, the parent's math mode) for work that does
 * not depend on the parent's data flow: z chains, kl_div.  Its stream runs `rank` levels above the device's LEAST
 * priority and always below a stream owned by mnf_ctx_create (the greatest), so side work fills in behind the
 * main stream instead of delaying it; rank the forks by how soon the main stream needs their results. */
int mnf_ctx_fork(mnf_ctx* parent, int32_t rank, mnf_ctx** out);
int mnf_ctx_destroy(mnf_ctx* ctx);
int mnf_ctx_set_stream(mnf_ctx* ctx, void* stream);
int mnf_ctx_get_stream(mnf_ctx* ctx, void** stream);
int mnf_ctx_sync(mnf_ctx* ctx);
/* Make all FUTURE work on ctx's stream wait for everything enqueued so far on other's stream
 * (event fork/join; lets kl_div(), which does not depend on x, run on a second context). */
int mnf_ctx_wait(mnf_ctx* ctx, mnf_ctx* other);
/* The same against a FOREIGN stream (the producer's stream of a tensor that enters through DLPack /
 * __cuda_array_interface__): a cudaStream_t, or 1 = legacy default stream, 2 = per-thread default
 * stream (the __cuda_array_interface__ v3 encoding).  Contexts own a non-blocking stream, so there is
 * no implicit ordering with a framework's streams. */
int mnf_ctx_wait_stream(mnf_ctx* ctx, void* stream);
int mnf_ctx_sm_count(mnf_ctx* ctx, int* out);
/* kernels launched by this context since creation (bench.py reports it as gpu_launches) */
int mnf_ctx_launch_count(mnf_ctx* ctx, int64_t* out);

/* Arithmetic of the paired mean/variance contractions (dense and conv forward):
 *  MNF_MATH_FP32   register-blocked FFMA kernels
 *  MNF_MATH_TF32X3 (default) tcgen05.mma with hi/lo operand splitting (3 MMAs per product, fp32
 *                  accumulation in TMEM): fp32-level accuracy (<= 1e-5 rel.) on tensor cores
 *  MNF_MATH_BF16   MNFDense contractions (N > 16) with bf16 operands, one tcgen05.mma kind::f16
 *                  per product, fp32 accumulation (BASELINE config "wide MNFDense 4096 -> 4096,
 *                  bf16 tensor-core path"): x*z and x^2 are formed in fp32 and rounded to nearest
 *                  bf16, as are W and sigma^2; mean and variance are then within 2^-7 relative of
 *                  the fp32 result term by term, ~1e-3 relative in L2 norm for K >= 64.  Conv
 *                  contractions of any length stay on the tensor cores (3xTF32; beyond 1280
 *                  terms the accumulation drift exceeds the 1e-5 bar, csrc/common.cuh).
 *                  Everything else (flows, KL, heads with N <= 16) runs as under MNF_MATH_TF32X3. */
enum { MNF_MATH_FP32 = 0, MNF_MATH_TF32X3 = 1, MNF_MATH_BF16 = 2 };
int mnf_ctx_set_math(mnf_ctx* ctx, int mode);
/* Inference with fixed parameters: while frozen != 0 the packed / pre-split / pre-scaled weight
 * blocks the tensor-core kernels consume are built once per (weight pointers, shape) and kept in
 * persistent device buffers instead of being rebuilt on every call.  The caller promises not to
 * change the parameter buffers while frozen; unfreezing (or freezing again) drops the blocks.
 * Results are bit-identical either way. */
int mnf_ctx_freeze_weights(mnf_ctx* ctx, int32_t frozen);

/* Timing probe for bench.py's roofline line: the (skip+1)-th launch of kernel family `kind`
 * (1 conv paired GEMM, 2 dense paired GEMM incl. its per-call operand packing, 3 flow step, 4 KL,
 * 5 Monte-Carlo expansion of a first conv layer; 0 disarms) is bracketed by
 * cudaEventRecord(ev_start/ev_stop) on the context's stream, then the probe disarms. */
int mnf_ctx_probe(mnf_ctx* ctx, int kind, int skip, void* ev_start, void* ev_stop);

int mnf_malloc(mnf_ctx* ctx, size_t bytes, void** out); /* stream-ordered pool */
int mnf_free(mnf_ctx* ctx, void* ptr);
int mnf_host_alloc(size_t bytes, void** out); /* pinned */
int mnf_host_free(void* ptr);
int mnf_memcpy_h2d(mnf_ctx* ctx, void* dst, const void* h_src, size_t bytes);
int mnf_memcpy_d2h(mnf_ctx* ctx, void* h_dst, const void* src, size_t bytes);
int mnf_memcpy_d2d(mnf_ctx* ctx, void* dst, const void* src, size_t bytes);
int mnf_memset(mnf_ctx* ctx, void* dst, int value, size_t bytes);

/* events for device-side timing on the context's stream */
int mnf_event_create(void** out);
int mnf_event_destroy(void* ev);
int mnf_event_record(mnf_ctx* ctx, void* ev);
int mnf_event_elapsed_ms(void* start, void* stop, float* out); /* synchronises on stop */

/* CUDA-graph capture of everything enqueued between begin/end on the context's stream */
int mnf_graph_begin(mnf_ctx* ctx);
/* Mark a context whose stream JOINS a capture (forked from the capturing context via mnf_ctx_wait):
 * while set, workspace growth / packed-weight cache misses on it fail with a clear message
 * ("run one eager step first") instead of issuing calls that invalidate the capture.
 * mnf_graph_begin / mnf_graph_end set and clear it on the capturing context itself. */
int mnf_ctx_set_capturing(mnf_ctx* ctx, int32_t capturing);
int mnf_graph_end(mnf_ctx* ctx, void** graph_exec_out);
int mnf_graph_launch(mnf_ctx* ctx, void* graph_exec);
int mnf_graph_destroy(void* graph_exec);

/* ---- DLPack (zero-copy exchange with TF / any producer) ------------------------------- */
typedef struct {
  void* data;       /* device pointer incl. byte_offset */
  int32_t ndim;
  int64_t shape[6];
  int32_t device_type; /* DLDeviceType; 2 = kDLCUDA */
  int32_t device_id;
  int32_t dtype_code, dtype_bits; /* 2,32 = float32 */
  int32_t contiguous;
} mnf_tensor_view;
/* dl_managed_tensor: the pointer held by a "dltensor" PyCapsule (DLManagedTensor*).
 * Borrowed view: the capsule keeps ownership. */
int mnf_dlpack_view(void* dl_managed_tensor, mnf_tensor_view* out);
/* Wrap device memory as a new DLManagedTensor*; when the consumer calls its deleter,
 * release(release_arg) is invoked (may be NULL). */
int mnf_dlpack_wrap(void* data, int32_t ndim, const int64_t* shape, int32_t device_id,
                    void (*release)(void*), void* release_arg, void** out_dl_managed_tensor);
/* Call the deleter of a DLManagedTensor* nobody consumed (capsule destructor path). */
int mnf_dlpack_delete(void* dl_managed_tensor);

/* ---- noise ---------------------------------------------------------------------------- */
typedef struct {
  uint64_t seed;         /* Philox key */
  uint64_t row_offset;   /* global index of local row 0 (shard invariance) */
  uint32_t stream;       /* sub-stream: distinguishes layers / draws */
  uint32_t reserved;
  const float* injected; /* if non-NULL: read the draw from here instead ([rows, outer, inner]) */
  const uint32_t* call_delta; /* if non-NULL: a DEVICE counter added to the call index (high key word) when the
                               * kernel runs -- how a captured CUDA graph draws fresh noise on every replay */
} mnf_noise;

/* Materialise exactly the values a kernel would draw for (seed, stream, row_offset):
 * replaces tf.random.normal (mnf_dense.py:94,141,167; mnf_conv.py:106,156,165,194). */
int mnf_philox_normal(mnf_ctx* ctx, uint64_t seed, uint32_t stream, uint64_t row_offset,
                      int64_t rows, int64_t outer, int64_t inner, float* out);
/* keras.backend.random_binomial(shape, p=0.5) (rnvp.py:36).  p == 0.5 (the RNVP mask) costs one
 * Philox BIT per element: element d of a row is bit (d & 31) of word ((d >> 5) & 3) of the block
 * with counter (d >> 7, 0, row, stream), and the draw is 1.0f when that bit is clear.  Any other
 * p spends one word per element: 1.0f where u01(word (d & 3) of block d >> 2) <= p. */
int mnf_philox_bernoulli(mnf_ctx* ctx, uint64_t seed, uint32_t stream, uint64_t row_offset,
                         int64_t rows, int64_t inner, float p, float* out);
/* raw 32-bit Philox words, for the bit-exact check against the oracle's integer Philox */
int mnf_philox_raw(mnf_ctx* ctx, uint64_t seed, uint32_t stream, uint64_t row_offset,
                   int64_t rows, int64_t outer, int64_t inner, uint32_t* out);

/* ---- flows: NormalizingFlow.forward (flows/__init__.py:22-29) --------------------------- */
enum { MNF_FLOW_IAF = 0, MNF_FLOW_RNVP = 1, MNF_FLOW_PLANAR = 2 };

/* One flow step.  D = flow dimension, H = hidden units of the first hidden layer.
 *  IAF    (maf.py:47-53, made.py:28-30,77-91): w1 [D,H], b1 [H] = first MaskedDense;
 *         ws/wt = LAST MaskedDense kernel [H_last,2D] columns [0,D) / [D,2D) (ld2 = 2D),
 *         bs/bt likewise halves of its bias; m1 [D,H], m2 [H_last,2D] optional masks (NULL = ones).
 *  RNVP   (rnvp.py:33-47): w1,b1 = first net Dense(H, activation); ws,bs = s Dense; wt,bt = t Dense
 *         (ld2 = D); bern = Bernoulli mask draw (injected or Philox).
 *  PLANAR (planar.py:22-36): w1 = dense kernel w [D], b1 = dense bias [1], wt = u [D].
 * General coupling networks (made.py:77-88 / rnvp.py:28-29 build one hidden layer per entry of
 * h_sizes): n_extra further hidden layers behind the first, layer l (1-based) mapping
 * extra_h[l-2]... i.e. extra_w[i] is [h_i, extra_h[i]] with h_0 = hidden, h_i = extra_h[i-1]; extra_m
 * are the MADE masks of those layers (NULL = ones).  activation: MNF_ACTV_DEFAULT = what the class
 * uses by default (MADE: ReLU, fixed; RNVP: tanh), or the RNVP constructor's `activation`.
 * One hidden layer with the default activation runs on the fused / tensor-core kernels; anything
 * else on the general kernels (csrc/flows_generic.cu: D <= 1024, hidden widths <= 256). */
#define MNF_FLOW_MAX_HIDDEN 4
enum { MNF_ACTV_DEFAULT = 0, MNF_ACTV_RELU = 1, MNF_ACTV_TANH = 2, MNF_ACTV_SIGMOID = 3, MNF_ACTV_LINEAR = 4 };
typedef struct {
  int32_t kind, parity, hidden, reserved;
  const float *w1, *b1, *ws, *bs, *wt, *bt;
  int64_t ld2;
  const float *m1, *m2;
  mnf_noise bern;
  int32_t n_extra, activation;
  int32_t extra_h[MNF_FLOW_MAX_HIDDEN - 1];
  int32_t reserved2;
  const float* extra_w[MNF_FLOW_MAX_HIDDEN - 1];
  const float* extra_b[MNF_FLOW_MAX_HIDDEN - 1];
  const float* extra_m[MNF_FLOW_MAX_HIDDEN - 1];
} mnf_flow_step;

/* sample_z + flow chain (mnf_dense.py:88-102, mnf_conv.py:100-114).
 *  z0 != NULL: start from z0 [B,D].  Otherwise z0 = q0_mean + exp(q0_log_var/2) * eps_z.
 *  states [T+1,B,D]: every intermediate state (the LIST the reference returns); the last
 *  slice is z_T.  log_dets [B] = sum of per-step log-determinants (may be NULL).
 *  saved (may be NULL): hidden activations kept for mnf_flow_backward, step after step, B x (sum of the
 *  step's hidden widths) floats each (nothing for planar steps). */
int mnf_flow_forward(mnf_ctx* ctx, int64_t B, int64_t D, int32_t T, const mnf_flow_step* h_steps,
                     const float* z0, const float* q0_mean, const float* q0_log_var,
                     const mnf_noise* eps_z, float* states, float* log_dets, float* saved);

/* Gradient of the chain (TF autodiff of the above).
 *  states [T+1,B,D], saved: what mnf_flow_forward wrote (saved must not be NULL).
 *  d_states [T+1,B,D] or NULL: upstream grads for every state; d_zT [B,D] or NULL: an
 *  additional upstream grad for the last state only; d_log_dets [B] (NULL = zeros).
 *  Outputs: d_z0 [B,D]; per-step parameter grads are ACCUMULATED (+=) into the buffers named
 *  by h_grad_steps (same slots as mnf_flow_step: w1,b1,ws,bs,wt,bt with ld2; planar: w1=dw,
 *  b1=db, wt=du), so the caller zeroes them.  IAF masks are applied to the kernel grads. */
int mnf_flow_backward(mnf_ctx* ctx, int64_t B, int64_t D, int32_t T, const mnf_flow_step* h_steps,
                      const float* states, const float* saved, const float* d_states, const float* d_zT,
                      const float* d_log_dets, float* d_z0, const mnf_flow_step* h_grad_steps);
/* MAF.forward == IAF.inverse, maf.py:33-45: the sequential decode of a MADE-based step, as the reference
 * spells it out (x = 0; z reversed when parity; for i < D: (s, t) = net(x); x[:, i] = (z[:, i] - t[:, i])
 * exp(-s[:, i]); log_det -= s[:, i]).  z, x [B,D]; log_dets [B] (may be NULL).  The single-pass direction
 * (MAF.inverse == IAF.forward) is mnf_flow_forward. */
int mnf_maf_decode(mnf_ctx* ctx, int64_t B, int64_t D, const mnf_flow_step* step, const float* z, float* x,
                   float* log_dets);
/* d q0_mean / d q0_log_var from d_z0 (App. A-2): accumulated (+=). */
int mnf_sample_z_backward(mnf_ctx* ctx, int64_t B, int64_t D, const float* d_z0,
                          const float* q0_log_var, const mnf_noise* eps_z,
                          float* d_q0_mean, float* d_q0_log_var);

/* ---- layers ----------------------------------------------------------------------------- */
/* MNFDense.call, mnf_dense.py:159-170:
 *   y = (x*z) @ mean_W + mean_b + sqrt(x^2 @ min(exp(log_std_W),max_std)^2
 *                                      + min(exp(log_var_b),max_std^2)) * eps
 *  z [B,K] or NULL (use_z=False).  sd_out [B,N] (may be NULL) receives sqrt(V) for backward. */
int mnf_dense_forward(mnf_ctx* ctx, int64_t B, int64_t K, int64_t N, const float* x, const float* z,
                      const float* mean_W, const float* log_std_W, const float* mean_b,
                      const float* log_var_b, float max_std, const mnf_noise* eps,
                      float* y, float* sd_out);
/* The same with the activation that follows the layer in the reference's models applied in the
 * kernel's epilogue (mnf_lenet.py:29-32: MNFDense, ReLU, MNFDense, Softmax; inference only --
 * sd_out then belongs to the pre-activation).  mnf_dense_can_fuse_act tells whether a shape /
 * math mode supports it (ReLU: tensor-core path and N <= 16; Softmax: N <= 16); otherwise
 * mnf_dense_forward_act returns MNF_ERR_UNSUPPORTED and the caller runs the stock layer. */
enum { MNF_ACT_NONE = 0, MNF_ACT_RELU = 1, MNF_ACT_SOFTMAX = 2 };
int mnf_dense_can_fuse_act(mnf_ctx* ctx, int64_t K, int64_t N, int32_t act, int32_t* ok);
int mnf_dense_forward_act(mnf_ctx* ctx, int64_t B, int64_t K, int64_t N, const float* x, const float* z,
                          const float* mean_W, const float* log_std_W, const float* mean_b,
                          const float* log_var_b, float max_std, const mnf_noise* eps, int32_t act,
                          float* y, float* sd_out);
/* App. A-1.  Outputs (all may be NULL to skip): dx [B,K], dz [B,K], d_mean_W, d_log_std_W [K,N],
 * d_mean_b, d_log_var_b [N] (overwritten, not accumulated). */
int mnf_dense_backward(mnf_ctx* ctx, int64_t B, int64_t K, int64_t N, const float* x, const float* z,
                       const float* mean_W, const float* log_std_W, const float* log_var_b,
                       float max_std, const mnf_noise* eps, const float* sd, const float* dy,
                       float* dx, float* dz, float* d_mean_W, float* d_log_std_W,
                       float* d_mean_b, float* d_log_var_b);

typedef struct {
  int64_t B, H, W, C;  /* input NHWC */
  int64_t kh, kw, F;   /* filters HWIO [kh,kw,C,F] */
  int64_t stride;      /* same in both spatial dims */
  int32_t same;        /* 0 = VALID, 1 = SAME (TF semantics) */
  int32_t fuse_relu_pool; /* forward only: apply ReLU + MaxPool2D(2x2, stride 2) in the epilogue
                             (mnf_lenet.py:23-24,26-27); y is then [B, Ho/2, Wo/2, F] */
} mnf_conv_shape;
int mnf_conv_out_hw(const mnf_conv_shape* s, int64_t* Ho, int64_t* Wo);
/* *ok = 1 when mnf_conv2d_forward can honour fuse_relu_pool for this shape under the context's
 * math mode (even Ho/Wo, filter count within one tile); otherwise call mnf_relu_maxpool2. */
int mnf_conv2d_can_fuse_pool(mnf_ctx* ctx, const mnf_conv_shape* s, int32_t* ok);

/* MNFConv2D.call, mnf_conv.py:182-197 (with the strides/padding the reference omits):
 *   y = (conv(x,mean_W)+mean_b) * z[b,f] + sqrt(conv(x^2, var_W) + var_b) * eps
 *  z [B,F] or NULL.  mean_out / sd_out [B,Ho,Wo,F] (may be NULL) saved for backward. */
int mnf_conv2d_forward(mnf_ctx* ctx, const mnf_conv_shape* s, const float* x, const float* z,
                       const float* mean_W, const float* log_std_W, const float* mean_b,
                       const float* log_var_b, float max_std, const mnf_noise* eps,
                       float* y, float* mean_out, float* sd_out);
/* App. A-1c. */
int mnf_conv2d_backward(mnf_ctx* ctx, const mnf_conv_shape* s, const float* x, const float* z,
                        const float* mean_W, const float* log_std_W, const float* log_var_b,
                        float max_std, const mnf_noise* eps, const float* mean_saved,
                        const float* sd, const float* dy, float* dx, float* dz,
                        float* d_mean_W, float* d_log_std_W, float* d_mean_b, float* d_log_var_b);

/* ---- KL --------------------------------------------------------------------------------- */
/* Parameters of one MNF layer as the KL needs them.  For MNFDense (mnf_dense.py:104-157):
 * rows = n_in = D, cols = n_out, z multiplies ROWS.  For MNFConv2D (mnf_conv.py:116-180):
 * rows = kh*kw*C, cols = F = D, z multiplies COLUMNS (conv = 1). */
typedef struct {
  int32_t conv;    /* 0 dense, 1 conv */
  int32_t use_z;
  int32_t r_all_states; /* 1: Gaussian r-term summed over all T+1 r-flow states (as written, App. B-3) */
  int32_t n_r_states;   /* T_r + 1 */
  int64_t rows, cols;
  const float *mean_W, *log_std_W, *mean_b, *log_var_b;
  const float *prior_var_r_p /*[rows]*/, *prior_var_r_p_bias /*[1]*/;
  const float *q0_log_var, *r0_mean, *r0_log_var, *r0_apvar; /* [D] */
  const float* z;        /* [D] final q-flow state of sample_z(1) */
  const float* ld_q;     /* [1] */
  const float* r_states; /* [n_r_states, D] states of flow_r.forward(z) */
  const float* ld_r;     /* [1] */
  mnf_noise eps_a;       /* dense: [cols]; conv: [rows] */
  mnf_noise eps_b;       /* conv only: scalar */
} mnf_kl_args;

typedef struct { /* gradient outputs of mnf_kl_backward; all overwritten; NULL = skip */
  float *mean_W, *log_std_W, *mean_b, *log_var_b, *prior_var_r_p, *prior_var_r_p_bias;
  float *q0_log_var, *r0_mean, *r0_log_var, *r0_apvar;
  float* d_z;        /* [D]  dKL/dz (direct terms; flow_r's contribution comes via d_r_states) */
  float* d_r_states; /* [n_r_states, D] */
} mnf_kl_grads;        /* note dKL/d(ld_q) = dKL/d(ld_r) = -scale (feed to mnf_flow_backward) */

/* kl_out [1] (device).  Single pass over the weights + one finishing block. */
int mnf_kl_forward(mnf_ctx* ctx, const mnf_kl_args* a, float* kl_out);
/* scale = upstream dLoss/dKL (host scalar, e.g. 1/len(X_train)). */
int mnf_kl_backward(mnf_ctx* ctx, const mnf_kl_args* a, float scale, mnf_kl_grads* g);

/* ---- stock layers between MNF layers (mnf_lenet.py:23-32) ------------------ */
/* ReLU followed by MaxPool2D(2x2, stride 2, padding SAME), NHWC.  argmax (may be NULL)
 * receives the int8 window index (0..3) of the max for the backward pass. */
int mnf_relu_maxpool2(mnf_ctx* ctx, int64_t B, int64_t H, int64_t W, int64_t C, const float* x,
                      float* y, int8_t* argmax);
int mnf_relu_maxpool2_backward(mnf_ctx* ctx, int64_t B, int64_t H, int64_t W, int64_t C,
                               const float* x, const int8_t* argmax, const float* dy, float* dx);
int mnf_relu(mnf_ctx* ctx, int64_t n, const float* x, float* y);
int mnf_relu_backward(mnf_ctx* ctx, int64_t n, const float* x, const float* dy, float* dx);
int mnf_softmax(mnf_ctx* ctx, int64_t B, int64_t N, const float* x, float* y);
/* mean categorical cross-entropy of softmax(logits) vs one-hot labels (lenet_mnist.py:53-57);
 * loss_out [1]; d_logits [B,N] (may be NULL) = (softmax - labels) / B_global. */
int mnf_softmax_xent(mnf_ctx* ctx, int64_t B, int64_t N, const float* logits, const float* labels,
                     float inv_global_batch, float* loss_out, float* d_logits);
/* loss_out [1] = mean squared error over the GLOBAL batch (inv_global_n = 1 / global element
 * count; bandgap_regression.py:57-69); dy (may be NULL) = 2 (y - target) * inv_global_n. */
int mnf_mse_loss(mnf_ctx* ctx, int64_t n, const float* y, const float* target, float inv_global_n,
                 float* loss_out, float* dy);
/* tf.keras.layers.BatchNormalization over the last axis of x [B,N] (mnf_feed_forward.py:20; Keras
 * defaults: momentum 0.99, epsilon 1e-3).  training != 0 (what model.fit runs, bandgap_regression.py:76):
 * the batch mean and biased batch variance normalise and moving_mean / moving_var move towards them
 * by (1 - momentum), in place; training == 0 (model(x) in the reference's GradientTape loop and at
 * prediction time): the moving statistics normalise.  save_mean / save_invstd [N] (NULL together to
 * skip) keep what the backward needs.  Backward: dx [B,N], d_gamma, d_beta [N] (each may be NULL;
 * overwritten); `training` must be the flag of the forward call.  Per-rank batch statistics under
 * data parallelism (no sync-BN), as Keras does per replica. */
int mnf_batchnorm_forward(mnf_ctx* ctx, int64_t B, int64_t N, const float* x, const float* gamma, const float* beta,
                          float* moving_mean, float* moving_var, float eps, float momentum, int32_t training, float* y,
                          float* save_mean, float* save_invstd);
int mnf_batchnorm_backward(mnf_ctx* ctx, int64_t B, int64_t N, const float* x, const float* gamma, const float* save_mean,
                           const float* save_invstd, int32_t training, const float* dy, float* dx, float* d_gamma,
                           float* d_beta);
/* y = a*x, generic axpby helpers */
int mnf_scale(mnf_ctx* ctx, int64_t n, float a, const float* x, float* y);
int mnf_axpy(mnf_ctx* ctx, int64_t n, float a, const float* x, float* y); /* y += a*x */
/* *counter += inc on the device (the replay counter of a captured step, see mnf_noise.call_delta) */
int mnf_u32_add(mnf_ctx* ctx, uint32_t* counter, uint32_t inc);
/* dst[(i * repeat + s) * row_elems + e] = src[i * row_elems + e]: the batch of a Monte-Carlo
 * prediction (every input row `repeat` times, np.repeat(x, repeat, axis=0)) formed on the
 * device, so only the distinct rows cross PCIe. */
int mnf_tile_rows(mnf_ctx* ctx, int64_t n_rows, int64_t row_elems, int64_t repeat, const float* src, float* dst);
/* Monte-Carlo predictive through a FIRST MNFConv2D layer: the batch repeats every
 * image n_rep times (lenet_mnist.py:84) and MNFConv2D.call (mnf_conv.py:182-197) applies z and eps
 * after the contractions, so conv(x, W_mu) + b_mu and sqrt(conv(x^2, sigma^2) + vb) are computed once
 * per DISTINCT image (mnf_conv2d_forward's mean_out / sd_out on the distinct rows) and expanded here:
 *   y[b] = mean[b / n_rep] * z[b] + sd[b / n_rep] * eps[b]      (+ ReLU and 2x2 max-pool when asked)
 * with the Philox counters of the full-batch kernels (global row, output pixel, filter quad).
 * mean, sd: [ceil(B / n_rep), Ho, Wo, F]; z: [B, F] or NULL; y: [B, Ho, Wo, F] or [B, Ho/2, Wo/2, F].
 * Needs F % 4 == 0 (and even Ho, Wo when pooling): mnf_conv2d_mc_can_expand tells. */
/* the per-image maps (a handful of images: one thread per output element) */
int mnf_conv2d_mc_maps(mnf_ctx* ctx, const mnf_conv_shape* s, const float* x, const float* mean_W, const float* log_std_W,
                       const float* mean_b, const float* log_var_b, float max_std, float* mean, float* sd);
int mnf_conv2d_mc_can_expand(mnf_ctx* ctx, int64_t Ho, int64_t Wo, int64_t F, int32_t fuse_relu_pool, int32_t* ok);
int mnf_conv2d_mc_expand(mnf_ctx* ctx, int64_t B, int64_t n_rep, int64_t Ho, int64_t Wo, int64_t F, const float* mean,
                         const float* sd, const float* z, const mnf_noise* eps, int32_t fuse_relu_pool, float* y);

/* The next MNFConv2D layer fed DIRECTLY from such a first layer (Monte-Carlo prediction, MNF-LeNet conv1 -> ReLU ->
 * MaxPool2D -> conv2, mnf_lenet.py:23-26): its input x = MaxPool2x2(ReLU(mnf_conv2d_mc_expand(...))) is described by
 * `src` instead of being materialised -- the tensor-core kernel's plane producers draw the first layer's noise
 * themselves (same Philox counters), so the [B, Ho/2, Wo/2, F] activation is neither written nor read and the
 * issue-bound noise generation runs beside the tensor-bound contraction on the same SMs.
 * s: the shape of THIS layer's input (H = src->Ho / 2, W = src->Wo / 2, C = src->F), fuse_relu_pool as in
 * mnf_conv2d_forward.  *used = 0: the shape / math mode is outside the fused kernel's domain (or src->eps is
 * injected) and nothing was launched -- expand, then call mnf_conv2d_forward. */
typedef struct {
  int64_t n_rep, Ho, Wo, F;  /* the first layer's pre-pool output map and the number of samples per distinct row */
  const float* mean;         /* [ceil(B / n_rep), Ho, Wo, F] from mnf_conv2d_mc_maps */
  const float* sd;
  const float* z;            /* [B, F] or NULL */
  mnf_noise eps;             /* the first layer's output noise (not injected) */
} mnf_mc_source;
int mnf_conv2d_forward_mc(mnf_ctx* ctx, const mnf_conv_shape* s, const mnf_mc_source* src, const float* z, const float* mean_W,
                          const float* log_std_W, const float* mean_b, const float* log_var_b, float max_std,
                          const mnf_noise* eps, float* y, int32_t* used);
/* ---- the data-parallel gradient all-reduce ---------------------------------------
 * The only collective of the design: where the reference applies its gradients
 * (notebooks/lenet_mnist.py:114-115, tape.gradient -> adam.apply_gradients), a data-parallel run sums
 * the flat fp32 gradient buffer over the ranks first.  One process per GPU; NCCL over NVLink, bound at
 * run time (dlopen of libnccl.so.2 -- the copy a host framework already loaded is re-used; override
 * with MNF_NCCL_LIB); without NCCL these calls return MNF_ERR_UNSUPPORTED, never a fallback.
 *   rank 0:    mnf_comm_unique_id(id)            -- host code ships the 128 bytes to the other ranks
 *   all ranks: mnf_comm_init(ctx, id, rank, world, &comm)          (collective; blocks until all joined)
 *   per step:  mnf_allreduce_grads(comm, ctx, grads, n)            in place, sum, on ctx's stream
 * The caller scales: likelihood gradients by 1/global batch (mnf_softmax_xent / mnf_mse_loss do), KL
 * gradients by 1/world, so that the plain sum is the gradient of the single-process loss. */
#define MNF_COMM_ID_BYTES 128
typedef struct mnf_comm mnf_comm;
int mnf_comm_nccl_version(int* out);
int mnf_comm_unique_id(uint8_t* h_id /* [MNF_COMM_ID_BYTES], host */);
int mnf_comm_init(mnf_ctx* ctx, const uint8_t* h_id, int rank, int world, mnf_comm** out);
int mnf_comm_info(mnf_comm* comm, int* rank, int* world);
int mnf_comm_destroy(mnf_comm* comm);
int mnf_allreduce_grads(mnf_comm* comm, mnf_ctx* ctx, float* grads, int64_t n);

/* Adam update on a flat parameter buffer (lenet_mnist.py:47,115). */
int mnf_adam_step(mnf_ctx* ctx, int64_t n, float* param, const float* grad, float* m, float* v,
                  float lr, float beta1, float beta2, float eps, int32_t step);
/* The same update with the step count kept on the device: d_state = {int32 step, float lr_t} (zero-initialised by
 * the caller); every call increments step and derives the bias-corrected rate from it, so a training step
 * captured into a CUDA graph (Trainer.capture) replays with the right rate. */
int mnf_adam_step_dev(mnf_ctx* ctx, int64_t n, float* param, const float* grad, float* m, float* v, float lr, float beta1,
                      float beta2, float eps, int32_t* d_state);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MNF_B200_H */
