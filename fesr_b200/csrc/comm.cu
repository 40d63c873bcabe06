// On-stream collectives of the sharded path: one NCCL communicator per process
// (= per GPU), the all-gather of the per-subdomain predictions and the gradient all-reduce issued on the
// caller's compute stream.  Replaces the reference's fan-out / fan-in through mp.Process + Manager().dict()
// (models/scheduler_gnn.py:254-291) and DistributedDataParallel's bucket all-reduce (:386).
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") -- the copy the host process has already loaded
// (PyTorch ships one) is reused, so libfesr.so itself has no link-time NCCL dependency and still loads on a
// machine without it (the CPU-side ABI test).  Only the handful of NCCL entry points used here are declared.
#include <dlfcn.h>

#include "common.cuh"

namespace fesr {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;        // NCCL_UNIQUE_ID_BYTES = 128 in every 2.x release
enum { ncclSuccess_ = 0 };
enum { ncclFloat32_ = 7 };                                   // ncclDataType_t::ncclFloat32
enum { ncclSum_ = 0 };                                       // ncclRedOp_t::ncclSum

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1;

static int load_nccl() {
  if (g_nccl.handle) return FESR_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);    // the copy the process already uses
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("fesr_comm: libnccl.so.2 not found (%s)", dlerror());
    return FESR_ECUDA;
  }
#define FESR_SYM(field, name)                                            \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);             \
  if (!g_nccl.field) {                                                   \
    set_error("fesr_comm: symbol %s missing from libnccl", name);        \
    return FESR_ECUDA;                                                   \
  }
  FESR_SYM(GetUniqueId, "ncclGetUniqueId")
  FESR_SYM(CommInitRank, "ncclCommInitRank")
  FESR_SYM(CommDestroy, "ncclCommDestroy")
  FESR_SYM(AllGather, "ncclAllGather")
  FESR_SYM(AllReduce, "ncclAllReduce")
  FESR_SYM(GetErrorString, "ncclGetErrorString")
#undef FESR_SYM
  g_nccl.handle = h;
  return FESR_OK;
}

#define FESR_NCCL(call)                                                                        \
  do {                                                                                         \
    int rc__ = (call);                                                                         \
    if (rc__ != ncclSuccess_) {                                                                \
      set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(rc__));        \
      return FESR_ECUDA;                                                                       \
    }                                                                                          \
  } while (0)

__global__ void scale_kernel(float* __restrict__ v, int64_t count, float s) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < count) v[i] *= s;
}

}  // namespace fesr

using namespace fesr;

extern "C" {

int fesr_comm_unique_id(void* host_id) {
  FESR_CHECK_ARG(host_id != nullptr, "host_id is NULL");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  FESR_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(host_id, &id, sizeof(id));
  return FESR_OK;
}

int fesr_comm_init(const void* host_id, int rank, int world) {
  FESR_CHECK_ARG(host_id != nullptr && world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
  FESR_CHECK_ARG(g_comm == nullptr, "fesr_comm_init: a communicator already exists (fesr_comm_destroy first)");
  int rc = load_nccl();
  if (rc) return rc;
  if ((rc = fesr_device_check())) return rc;
  ncclUniqueId id;
  memcpy(&id, host_id, sizeof(id));
  FESR_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
  g_rank = rank;
  g_world = world;
  return FESR_OK;
}

int fesr_comm_destroy(void) {
  if (g_comm) {
    FESR_NCCL(g_nccl.CommDestroy(g_comm));
    g_comm = nullptr;
  }
  g_rank = 0;
  g_world = 1;
  return FESR_OK;
}

int fesr_comm_rank(void) { return g_comm ? g_rank : -1; }
int fesr_comm_world(void) { return g_comm ? g_world : 0; }

int fesr_allgatherv_pred(float* slots, int64_t slot_elems, void* stream) {
  FESR_CHECK_ARG(g_comm != nullptr, "fesr_allgatherv_pred: no communicator (fesr_comm_init)");
  FESR_CHECK_ARG(slots != nullptr && slot_elems > 0, "bad slot buffer");
  // in place: rank r's send block IS slot r of the receive buffer (NCCL's in-place all-gather contract)
  FESR_NCCL(g_nccl.AllGather(slots + (size_t)g_rank * slot_elems, slots, (size_t)slot_elems, ncclFloat32_, g_comm,
                             as_stream(stream)));
  count_launch();
  return FESR_OK;
}

int fesr_allreduce_grads(float* flat, int64_t count, void* stream) {
  FESR_CHECK_ARG(g_comm != nullptr, "fesr_allreduce_grads: no communicator (fesr_comm_init)");
  FESR_CHECK_ARG(flat != nullptr && count > 0, "bad gradient buffer");
  cudaStream_t s = as_stream(stream);
  FESR_NCCL(g_nccl.AllReduce(flat, flat, (size_t)count, ncclFloat32_, ncclSum_, g_comm, s));
  count_launch();
  // DDP semantics: the mean over the ranks (sum, then one multiply -- the same arithmetic for every world size)
  scale_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, s>>>(flat, count, 1.0f / (float)g_world);
  FESR_LAUNCH_CHECK();
  return FESR_OK;
}

}  // extern "C"
