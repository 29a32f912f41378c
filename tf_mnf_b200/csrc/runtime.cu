// Runtime plumbing of libmnf_b200: context, memory, events, CUDA graphs, DLPack views.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void mnf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local const char* g_path = "";
void mnf_note_path(const char* name) { g_path = name; }

void* mnf_workspace(mnf_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return ctx->ws;
  if (ctx->capturing) {
    mnf_set_error("workspace growth (%zu bytes) during CUDA-graph capture; run one eager step first", bytes);
    return nullptr;
  }
  cudaError_t e;
  if (ctx->ws) {
    e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaFree(ctx->ws);
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    if (e != cudaSuccess) {
      mnf_set_error("workspace free failed: %s", cudaGetErrorString(e));
      return nullptr;
    }
  }
  size_t want = bytes + bytes / 4 + (1 << 20);
  e = cudaMalloc(&ctx->ws, want);
  if (e != cudaSuccess) {
    mnf_set_error("workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    ctx->ws = nullptr;
    return nullptr;
  }
  ctx->ws_bytes = want;
  return ctx->ws;
}

void* mnf_packed(mnf_ctx* ctx, uint64_t k0, uint64_t k1, uint64_t k2, uint64_t k3, size_t bytes, bool* fresh) {
  *fresh = true;
  if (!ctx->frozen) return mnf_workspace(ctx, bytes);
  for (int i = 0; i < ctx->packed_n; ++i) {
    mnf_packed_entry& e = ctx->packed[i];
    if (e.key[0] == k0 && e.key[1] == k1 && e.key[2] == k2 && e.key[3] == k3 && e.bytes == bytes) {
      *fresh = false;
      return e.ptr;
    }
  }
  if (ctx->capturing) {
    mnf_set_error("packed-weight cache miss during CUDA-graph capture; run one eager step first");
    return nullptr;
  }
  if (ctx->packed_n == ctx->packed_cap) {
    const int cap = ctx->packed_cap ? ctx->packed_cap * 2 : 32;
    mnf_packed_entry* p = (mnf_packed_entry*)realloc(ctx->packed, sizeof(mnf_packed_entry) * cap);
    if (!p) {
      mnf_set_error("packed-weight cache: out of host memory");
      return nullptr;
    }
    ctx->packed = p;
    ctx->packed_cap = cap;
  }
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, bytes);
  if (e != cudaSuccess) {
    mnf_set_error("packed-weight cache cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  mnf_packed_entry& n = ctx->packed[ctx->packed_n++];
  n.key[0] = k0; n.key[1] = k1; n.key[2] = k2; n.key[3] = k3;
  n.ptr = d;
  n.bytes = bytes;
  return d;
}

int* mnf_geom_table(mnf_ctx* ctx, uint64_t key, const int* h_data, size_t n) {
  uint64_t sum = 0xCBF29CE484222325ull;  // FNV-1a over the table: two shapes never share an entry by accident
  for (size_t i = 0; i < n; ++i) sum = (sum ^ (uint32_t)h_data[i]) * 0x100000001B3ull;
  for (int i = 0; i < ctx->geom_n; ++i) {
    mnf_packed_entry& e = ctx->geom[i];
    if (e.key[0] == key && e.key[1] == sum && e.bytes == n * sizeof(int)) return (int*)e.ptr;
  }
  if (ctx->capturing) {
    mnf_set_error("geometry-table cache miss during CUDA-graph capture; run one eager step first");
    return nullptr;
  }
  if (ctx->geom_n == ctx->geom_cap) {
    const int cap = ctx->geom_cap ? ctx->geom_cap * 2 : 16;
    mnf_packed_entry* p = (mnf_packed_entry*)realloc(ctx->geom, sizeof(mnf_packed_entry) * cap);
    if (!p) {
      mnf_set_error("geometry-table cache: out of host memory");
      return nullptr;
    }
    ctx->geom = p;
    ctx->geom_cap = cap;
  }
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, n * sizeof(int) + 16);
  if (e == cudaSuccess) e = cudaMemcpy(d, h_data, n * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    mnf_set_error("geometry-table upload (%zu ints) failed: %s", n, cudaGetErrorString(e));
    return nullptr;
  }
  mnf_packed_entry& ne = ctx->geom[ctx->geom_n++];
  ne.key[0] = key; ne.key[1] = sum; ne.key[2] = ne.key[3] = 0;
  ne.ptr = d;
  ne.bytes = n * sizeof(int);
  return (int*)d;
}

static void mnf_packed_clear(mnf_ctx* ctx) {
  if (ctx->packed_n) cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < ctx->packed_n; ++i) cudaFree(ctx->packed[i].ptr);
  ctx->packed_n = 0;
}

extern "C" {

int mnf_ctx_freeze_weights(mnf_ctx* ctx, int32_t frozen) {
  MNF_CHECK_ARG(ctx, "mnf_ctx_freeze_weights: ctx is NULL");
  MNF_CUDA(cudaSetDevice(ctx->device));
  if (!frozen || !ctx->frozen) mnf_packed_clear(ctx);  // (re)freezing starts from the current weights
  ctx->frozen = frozen != 0;
  return MNF_OK;
}

int mnf_version(void) { return MNF_B200_VERSION; }
const char* mnf_last_error(void) { return g_err; }
const char* mnf_last_path(void) { return g_path; }

int mnf_device_count(int* out) {
  MNF_CHECK_ARG(out, "mnf_device_count: out is NULL");
  MNF_CUDA(cudaGetDeviceCount(out));
  return MNF_OK;
}

static int ctx_create(int device, void* stream, int rank, mnf_ctx** out) {  // rank < 0: the greatest priority; else `rank` levels above the least
  MNF_CHECK_ARG(out, "mnf_ctx_create: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    mnf_set_error("mnf_ctx_create: no CUDA device available (%s); libmnf_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return MNF_ERR_CUDA;
  }
  MNF_CHECK_ARG(device >= 0 && device < n, "mnf_ctx_create: device %d out of range [0,%d)", device, n);
  // One device per process (one process per GPU): the entry points launch on the
  // CURRENT device and the kernels' function attributes are set once per process, so a second
  // device in the same process is refused instead of failing later with an opaque CUDA error.
  static int first_device = -1;
  if (first_device >= 0 && device != first_device) {
    mnf_set_error("mnf_ctx_create: this process already drives device %d; libmnf_b200 uses one process per GPU "
                  "(asked for device %d)", first_device, device);
    return MNF_ERR_UNSUPPORTED;
  }
  MNF_CUDA(cudaSetDevice(device));
  first_device = device;
  mnf_ctx* c = (mnf_ctx*)calloc(1, sizeof(mnf_ctx));
  c->device = device;
  c->math_mode = MNF_MATH_TF32X3;  // tensor cores with fp32-level accuracy: what a drop-in user gets
  if (stream) {
    c->stream = (cudaStream_t)stream;
    c->owns_stream = false;
  } else {
    // An owned stream of mnf_ctx_create runs at the device's greatest priority, a fork's (mnf_ctx_fork) below it:
    // the block scheduler then hands free SMs to the x-dependent main stream first and the side work (z chains,
    // KL terms) fills in behind it -- the main stream's kernels run at their isolated durations.  Forks are
    // ranked among themselves (`rank` levels above the least priority): the z chain the main stream needs NEXT
    // must not queue behind the ones it needs later (a priority inversion otherwise: the main stream waits).
    int least = 0, greatest = 0;
    MNF_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    int prio = greatest;
    if (rank >= 0) {  // (numerically smaller = more urgent; a fork stays below the main streams)
      prio = least - rank;
      if (prio <= greatest) prio = greatest + (least > greatest ? 1 : 0);
    }
    MNF_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio));
    c->owns_stream = true;
  }
  MNF_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  // keep freed blocks cached in the stream-ordered pool instead of returning them to the OS
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = c;
  return MNF_OK;
}

int mnf_ctx_create(int device, void* stream, mnf_ctx** out) { return ctx_create(device, stream, -1, out); }

int mnf_ctx_fork(mnf_ctx* parent, int32_t rank, mnf_ctx** out) {
  MNF_CHECK_ARG(parent && out && rank >= 0, "mnf_ctx_fork: bad argument");
  const int rc = ctx_create(parent->device, nullptr, rank, out);
  if (rc) return rc;
  (*out)->math_mode = parent->math_mode;
  return MNF_OK;
}

int mnf_ctx_destroy(mnf_ctx* ctx) {
  if (!ctx) return MNF_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  mnf_packed_clear(ctx);
  free(ctx->packed);
  for (int i = 0; i < ctx->geom_n; ++i) cudaFree(ctx->geom[i].ptr);
  free(ctx->geom);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  free(ctx);
  return MNF_OK;
}

int mnf_ctx_set_stream(mnf_ctx* ctx, void* stream) {
  MNF_CHECK_ARG(ctx, "ctx is NULL");
  if (ctx->owns_stream && stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
    ctx->owns_stream = false;
  }
  if (stream) ctx->stream = (cudaStream_t)stream;
  return MNF_OK;
}

int mnf_ctx_get_stream(mnf_ctx* ctx, void** stream) {
  MNF_CHECK_ARG(ctx && stream, "ctx/stream is NULL");
  *stream = (void*)ctx->stream;
  return MNF_OK;
}

int mnf_ctx_sync(mnf_ctx* ctx) {
  MNF_CHECK_ARG(ctx, "ctx is NULL");
  MNF_CUDA(cudaStreamSynchronize(ctx->stream));
  return MNF_OK;
}

int mnf_ctx_wait(mnf_ctx* ctx, mnf_ctx* other) {
  MNF_CHECK_ARG(ctx && other, "mnf_ctx_wait: NULL context");
  if (ctx == other || ctx->stream == other->stream) return MNF_OK;
  cudaEvent_t ev;
  MNF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, other->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev, 0);
  cudaEventDestroy(ev);  // released once the recorded work completes
  if (e != cudaSuccess) {
    mnf_set_error("mnf_ctx_wait failed: %s", cudaGetErrorString(e));
    return MNF_ERR_CUDA;
  }
  return MNF_OK;
}

int mnf_ctx_wait_stream(mnf_ctx* ctx, void* stream) {
  MNF_CHECK_ARG(ctx, "mnf_ctx_wait_stream: ctx is NULL");
  // __cuda_array_interface__ v3 stream encoding: 1 = legacy default stream, 2 = per-thread default
  cudaStream_t s = stream == (void*)1 ? cudaStreamLegacy : stream == (void*)2 ? cudaStreamPerThread : (cudaStream_t)stream;
  if (s == ctx->stream) return MNF_OK;
  cudaEvent_t ev;
  MNF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, s);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev, 0);
  cudaEventDestroy(ev);
  if (e != cudaSuccess) {
    mnf_set_error("mnf_ctx_wait_stream failed: %s", cudaGetErrorString(e));
    return MNF_ERR_CUDA;
  }
  return MNF_OK;
}

int mnf_ctx_sm_count(mnf_ctx* ctx, int* out) {
  MNF_CHECK_ARG(ctx && out, "ctx/out is NULL");
  *out = ctx->sm_count;
  return MNF_OK;
}

int mnf_ctx_launch_count(mnf_ctx* ctx, int64_t* out) {
  MNF_CHECK_ARG(ctx && out, "ctx/out is NULL");
  *out = ctx->launches;
  return MNF_OK;
}

int mnf_ctx_set_math(mnf_ctx* ctx, int mode) {
  MNF_CHECK_ARG(ctx, "ctx is NULL");
  MNF_CHECK_ARG(mode == MNF_MATH_FP32 || mode == MNF_MATH_TF32X3 || mode == MNF_MATH_BF16, "mnf_ctx_set_math: unknown mode %d", mode);
  ctx->math_mode = mode;
  return MNF_OK;
}

int mnf_ctx_probe(mnf_ctx* ctx, int kind, int skip, void* ev_start, void* ev_stop) {
  MNF_CHECK_ARG(ctx, "ctx is NULL");
  MNF_CHECK_ARG(kind == MNF_PROBE_NONE || (ev_start && ev_stop && skip >= 0), "mnf_ctx_probe: bad arguments");
  ctx->probe_kind = kind;
  ctx->probe_skip = skip;
  ctx->probe_start = (cudaEvent_t)ev_start;
  ctx->probe_stop = (cudaEvent_t)ev_stop;
  return MNF_OK;
}

int mnf_malloc(mnf_ctx* ctx, size_t bytes, void** out) {
  MNF_CHECK_ARG(ctx && out, "ctx/out is NULL");
  if (bytes == 0) bytes = 16;
  MNF_CUDA(cudaMallocAsync(out, bytes, ctx->stream));
  return MNF_OK;
}

int mnf_free(mnf_ctx* ctx, void* ptr) {
  MNF_CHECK_ARG(ctx, "ctx is NULL");
  if (ptr) MNF_CUDA(cudaFreeAsync(ptr, ctx->stream));
  return MNF_OK;
}

int mnf_host_alloc(size_t bytes, void** out) {
  MNF_CHECK_ARG(out, "out is NULL");
  MNF_CUDA(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault));
  return MNF_OK;
}

int mnf_host_free(void* ptr) {
  if (ptr) MNF_CUDA(cudaFreeHost(ptr));
  return MNF_OK;
}

int mnf_memcpy_h2d(mnf_ctx* ctx, void* dst, const void* src, size_t bytes) {
  MNF_CHECK_ARG(ctx && (bytes == 0 || (dst && src)), "mnf_memcpy_h2d: NULL argument");
  if (bytes) MNF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return MNF_OK;
}

int mnf_memcpy_d2h(mnf_ctx* ctx, void* dst, const void* src, size_t bytes) {
  MNF_CHECK_ARG(ctx && (bytes == 0 || (dst && src)), "mnf_memcpy_d2h: NULL argument");
  if (bytes) MNF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return MNF_OK;
}

int mnf_memcpy_d2d(mnf_ctx* ctx, void* dst, const void* src, size_t bytes) {
  MNF_CHECK_ARG(ctx && (bytes == 0 || (dst && src)), "mnf_memcpy_d2d: NULL argument");
  if (bytes) MNF_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return MNF_OK;
}

int mnf_memset(mnf_ctx* ctx, void* dst, int value, size_t bytes) {
  MNF_CHECK_ARG(ctx && (bytes == 0 || dst), "mnf_memset: NULL argument");
  if (bytes) MNF_CUDA(cudaMemsetAsync(dst, value, bytes, ctx->stream));
  return MNF_OK;
}

int mnf_event_create(void** out) {
  MNF_CHECK_ARG(out, "out is NULL");
  cudaEvent_t ev;
  MNF_CUDA(cudaEventCreate(&ev));
  *out = (void*)ev;
  return MNF_OK;
}

int mnf_event_destroy(void* ev) {
  if (ev) MNF_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return MNF_OK;
}

int mnf_event_record(mnf_ctx* ctx, void* ev) {
  MNF_CHECK_ARG(ctx && ev, "ctx/event is NULL");
  MNF_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx->stream));
  return MNF_OK;
}

int mnf_event_elapsed_ms(void* start, void* stop, float* out) {
  MNF_CHECK_ARG(start && stop && out, "NULL argument");
  // a failure here (an event that was never recorded: a probe whose kernel family did not run) must not stay
  // behind as the "last error" that the next launch check then reports as its own
  cudaError_t e = cudaEventSynchronize((cudaEvent_t)stop);
  if (e == cudaSuccess) e = cudaEventElapsedTime(out, (cudaEvent_t)start, (cudaEvent_t)stop);
  if (e != cudaSuccess) {
    cudaGetLastError();
    mnf_set_error("mnf_event_elapsed_ms: %s", cudaGetErrorString(e));
    return MNF_ERR_CUDA;
  }
  return MNF_OK;
}

int mnf_graph_begin(mnf_ctx* ctx) {
  MNF_CHECK_ARG(ctx, "ctx is NULL");
  MNF_CHECK_ARG(!ctx->capturing, "mnf_graph_begin: already capturing");
  MNF_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  ctx->capturing = true;
  return MNF_OK;
}

int mnf_ctx_set_capturing(mnf_ctx* ctx, int32_t capturing) {
  MNF_CHECK_ARG(ctx, "ctx is NULL");
  ctx->capturing = capturing != 0;
  return MNF_OK;
}

int mnf_graph_end(mnf_ctx* ctx, void** graph_exec_out) {
  MNF_CHECK_ARG(ctx && graph_exec_out, "NULL argument");
  MNF_CHECK_ARG(ctx->capturing, "mnf_graph_end: not capturing");
  ctx->capturing = false;
  cudaGraph_t g;
  MNF_CUDA(cudaStreamEndCapture(ctx->stream, &g));
  cudaGraphExec_t ge;
  // the kernel nodes carry the priority of the stream they were captured on (main stream: greatest, forks: least,
  // mnf_ctx_fork); without this flag a replay runs every node at the launching stream's priority
  cudaError_t e = cudaGraphInstantiate(&ge, g, cudaGraphInstantiateFlagUseNodePriority);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) {
    mnf_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    return MNF_ERR_CUDA;
  }
  *graph_exec_out = (void*)ge;
  return MNF_OK;
}

int mnf_graph_launch(mnf_ctx* ctx, void* graph_exec) {
  MNF_CHECK_ARG(ctx && graph_exec, "NULL argument");
  MNF_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, ctx->stream));
  return MNF_OK;
}

int mnf_graph_destroy(void* graph_exec) {
  if (graph_exec) MNF_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return MNF_OK;
}

// ------------------------------------------------------------------------------ DLPack
// Minimal DLPack v0.x structs (ABI-stable subset used by tf.experimental.dlpack / torch).
typedef struct { int32_t device_type; int32_t device_id; } DLDevice_;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType_;
typedef struct {
  void* data; DLDevice_ device; int32_t ndim; DLDataType_ dtype;
  int64_t* shape; int64_t* strides; uint64_t byte_offset;
} DLTensor_;
typedef struct DLManagedTensor_ {
  DLTensor_ dl_tensor; void* manager_ctx; void (*deleter)(struct DLManagedTensor_*);
} DLManagedTensor_;

int mnf_dlpack_view(void* p, mnf_tensor_view* out) {
  MNF_CHECK_ARG(p && out, "mnf_dlpack_view: NULL argument");
  const DLTensor_* t = &((DLManagedTensor_*)p)->dl_tensor;
  MNF_CHECK_ARG(t->ndim >= 0 && t->ndim <= 6, "mnf_dlpack_view: ndim %d not in [0,6]", t->ndim);
  memset(out, 0, sizeof(*out));
  out->data = (char*)t->data + t->byte_offset;
  out->ndim = t->ndim;
  out->device_type = t->device.device_type;
  out->device_id = t->device.device_id;
  out->dtype_code = t->dtype.code;
  out->dtype_bits = t->dtype.bits;
  int64_t expect = 1;
  out->contiguous = 1;
  for (int i = t->ndim - 1; i >= 0; --i) {
    out->shape[i] = t->shape[i];
    if (t->strides && t->shape[i] != 1 && t->strides[i] != expect) out->contiguous = 0;
    expect *= t->shape[i];
  }
  if (t->dtype.lanes != 1) out->contiguous = 0;
  if (out->device_type != 2 /*kDLCUDA*/ && out->device_type != 13 /*kDLCUDAManaged*/) {
    mnf_set_error("tensor lives on DLPack device type %d, not CUDA; libmnf_b200 has no CPU path",
                  out->device_type);
    return MNF_ERR_NOT_CUDA;
  }
  return MNF_OK;
}

struct WrapCtx {
  int64_t shape[6];
  void (*release)(void*);
  void* arg;
};

static void wrap_deleter(DLManagedTensor_* m) {
  WrapCtx* w = (WrapCtx*)m->manager_ctx;
  if (w->release) w->release(w->arg);
  free(w);
  free(m);
}

int mnf_dlpack_wrap(void* data, int32_t ndim, const int64_t* shape, int32_t device_id,
                    void (*release)(void*), void* release_arg, void** out) {
  MNF_CHECK_ARG(out && (ndim == 0 || shape) && ndim >= 0 && ndim <= 6, "mnf_dlpack_wrap: bad arguments");
  DLManagedTensor_* m = (DLManagedTensor_*)calloc(1, sizeof(DLManagedTensor_));
  WrapCtx* w = (WrapCtx*)calloc(1, sizeof(WrapCtx));
  for (int i = 0; i < ndim; ++i) w->shape[i] = shape[i];
  w->release = release;
  w->arg = release_arg;
  m->dl_tensor.data = data;
  m->dl_tensor.device.device_type = 2;
  m->dl_tensor.device.device_id = device_id;
  m->dl_tensor.ndim = ndim;
  m->dl_tensor.dtype.code = 2;
  m->dl_tensor.dtype.bits = 32;
  m->dl_tensor.dtype.lanes = 1;
  m->dl_tensor.shape = w->shape;
  m->dl_tensor.strides = nullptr;
  m->dl_tensor.byte_offset = 0;
  m->manager_ctx = w;
  m->deleter = wrap_deleter;
  *out = m;
  return MNF_OK;
}

// Run the deleter of a DLManagedTensor that no consumer took (unconsumed capsule).
int mnf_dlpack_delete(void* p) {
  if (!p) return MNF_OK;
  DLManagedTensor_* m = (DLManagedTensor_*)p;
  if (m->deleter) m->deleter(m);
  return MNF_OK;
}

}  // extern "C"
