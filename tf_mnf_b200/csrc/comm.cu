// The one collective of the design: the data-parallel gradient all-reduce of the
// training configs, where the reference applies its gradients (notebooks/lenet_mnist.py:114-115:
// tape.gradient -> adam.apply_gradients).  One ncclAllReduce(sum) over the Trainer's flat fp32 gradient
// buffer, in place, enqueued on the context's stream (no host synchronisation).
//
// NCCL is bound at run time (dlopen) rather than at link time: the library stays loadable on a box
// without NCCL, and inside a process that already carries an NCCL (PyTorch bundles its own
// libnccl.so.2) the SAME copy is used -- dlopen by soname returns the loaded object.  Only the few
// entry points below are needed, so their prototypes are restated here instead of including nccl.h.
#include <dlfcn.h>

#include <cstdlib>

#include "common.cuh"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return MNF_OK;
  const char* names[] = {getenv("MNF_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names)
    if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!h) {
    mnf_set_error("mnf_comm: cannot load NCCL (libnccl.so.2; set MNF_NCCL_LIB): %s", dlerror());
    return MNF_ERR_UNSUPPORTED;
  }
  NcclApi a;
  a.handle = h;
  *(void**)&a.GetVersion = dlsym(h, "ncclGetVersion");
  *(void**)&a.GetUniqueId = dlsym(h, "ncclGetUniqueId");
  *(void**)&a.CommInitRank = dlsym(h, "ncclCommInitRank");
  *(void**)&a.CommDestroy = dlsym(h, "ncclCommDestroy");
  *(void**)&a.AllReduce = dlsym(h, "ncclAllReduce");
  *(void**)&a.GetErrorString = dlsym(h, "ncclGetErrorString");
  if (!a.GetVersion || !a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString) {
    mnf_set_error("mnf_comm: the loaded NCCL lacks a required entry point");
    dlclose(h);
    return MNF_ERR_UNSUPPORTED;
  }
  g_nccl = a;
  return MNF_OK;
}
}  // namespace

struct mnf_comm {
  ncclComm_t comm;
  int rank, world, device;
};

#define MNF_NCCL(call)                                                                   \
  do {                                                                                   \
    ncclResult_t r__ = (call);                                                           \
    if (r__ != 0) {                                                                      \
      mnf_set_error("%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
      return MNF_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

int mnf_comm_nccl_version(int* out) {
  MNF_CHECK_ARG(out, "mnf_comm_nccl_version: out is NULL");
  int rc = load_nccl();
  if (rc != MNF_OK) return rc;
  MNF_NCCL(g_nccl.GetVersion(out));
  return MNF_OK;
}

int mnf_comm_unique_id(uint8_t* h_id) {
  MNF_CHECK_ARG(h_id, "mnf_comm_unique_id: h_id is NULL");
  static_assert(sizeof(ncclUniqueId) == MNF_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  int rc = load_nccl();
  if (rc != MNF_OK) return rc;
  ncclUniqueId id;
  MNF_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(h_id, id.internal, MNF_COMM_ID_BYTES);
  return MNF_OK;
}

int mnf_comm_init(mnf_ctx* ctx, const uint8_t* h_id, int rank, int world, mnf_comm** out) {
  MNF_CHECK_ARG(ctx && h_id && out, "mnf_comm_init: NULL argument");
  MNF_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "mnf_comm_init: rank %d not in [0,%d)", rank, world);
  int rc = load_nccl();
  if (rc != MNF_OK) return rc;
  MNF_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(id.internal, h_id, MNF_COMM_ID_BYTES);
  ncclComm_t comm = nullptr;
  MNF_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  mnf_comm* c = (mnf_comm*)calloc(1, sizeof(mnf_comm));
  c->comm = comm; c->rank = rank; c->world = world; c->device = ctx->device;
  *out = c;
  return MNF_OK;
}

int mnf_comm_destroy(mnf_comm* c) {
  if (!c) return MNF_OK;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  free(c);
  return MNF_OK;
}

int mnf_comm_info(mnf_comm* c, int* rank, int* world) {
  MNF_CHECK_ARG(c, "mnf_comm_info: comm is NULL");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  return MNF_OK;
}

int mnf_allreduce_grads(mnf_comm* c, mnf_ctx* ctx, float* grads, int64_t n) {
  MNF_CHECK_ARG(c && ctx && (grads || n == 0) && n >= 0, "mnf_allreduce_grads: bad arguments");
  MNF_CHECK_ARG(c->device == ctx->device, "mnf_allreduce_grads: communicator lives on device %d, context on %d", c->device,
                ctx->device);
  if (n == 0 || c->world == 1) return MNF_OK;
  MNF_NCCL(g_nccl.AllReduce(grads, grads, (size_t)n, kNcclFloat32, kNcclSum, c->comm, ctx->stream));
  ctx->launches++;  // NCCL's reduction kernel
  return MNF_OK;
}
