// comm.cu — data-parallel gradient exchange: NCCL all-reduce of the flat fp32 gradient bucket over NVLink/NVSwitch.
// The reference is single-process (no NCCL/MPI anywhere); this exchange step is defined by BASELINE's
// north_star. One process per GPU; the bucket all-reduce runs on a side stream ordered after the compute stream
// (event), so it can overlap whatever the compute stream does next; pdn_allreduce_wait() makes the compute stream wait
// for it (before the fused Adam kernel, which also folds in the 1/world scale). Small per-feature statistics of
// batch-coupled norms are reduced inline on the compute stream (pdn_allreduce_sum_f32_inline).
// NCCL is resolved with dlopen at first use, so libpdn_b200.so loads on hosts without it.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

namespace pdn {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;  // optional (NCCL >= 2.18)
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi     g_nccl;
static ncclComm_t  g_comm = nullptr;         // compute stream: small inline reductions (batch statistics of coupled norms)
static ncclComm_t  g_comm_bucket = nullptr;  // side stream: gradient-bucket all-reduces (its own communicator, so the two streams
                                             // never serialise on one NCCL work queue while backward is still running)
static int         g_world = 1, g_rank = 0;
static cudaEvent_t g_ev_compute = nullptr, g_ev_comm = nullptr;

static int load_nccl() {
  if (g_nccl.handle) return 0;
  const char* names[] = {getenv("PDN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) {
    set_error("cannot dlopen libnccl.so.2 (%s); set PDN_NCCL_LIB", dlerror());
    return PDN_ERR_NCCL;
  }
#define PDN_SYM(field, name)                                                  \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name);                      \
  if (!g_nccl.field) { set_error("libnccl lacks %s", name); return PDN_ERR_NCCL; }
  PDN_SYM(GetUniqueId, "ncclGetUniqueId")
  PDN_SYM(CommInitRank, "ncclCommInitRank")
  PDN_SYM(AllReduce, "ncclAllReduce")
  PDN_SYM(CommDestroy, "ncclCommDestroy")
  PDN_SYM(GetErrorString, "ncclGetErrorString")
#undef PDN_SYM
  *(void**)(&g_nccl.CommSplit) = dlsym(g_nccl.handle, "ncclCommSplit");
  *(void**)(&g_nccl.CommCount) = dlsym(g_nccl.handle, "ncclCommCount");
  *(void**)(&g_nccl.CommUserRank) = dlsym(g_nccl.handle, "ncclCommUserRank");
  return 0;
}

#define PDN_NCCL(expr)                                                                      \
  do {                                                                                      \
    ncclResult_t _r = (expr);                                                               \
    if (_r != ncclSuccess) {                                                                \
      set_error("NCCL error %d (%s) in `%s`", (int)_r, g_nccl.GetErrorString(_r), #expr);   \
      return PDN_ERR_NCCL;                                                                  \
    }                                                                                       \
  } while (0)

}  // namespace pdn

using namespace pdn;

extern "C" {

int pdn_nccl_unique_id(char* id128) {
  PDN_TRY(load_nccl());
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  PDN_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return 0;
}

int pdn_nccl_init(int rank, int world, const char* id128) {
  PDN_TRY(ensure_init());
  PDN_TRY(load_nccl());
  PDN_CHECK(!g_comm, "NCCL communicator already initialised");
  PDN_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  PDN_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
  g_comm_bucket = g_comm;
  if (g_nccl.CommSplit && world > 1 && !getenv("PDN_NCCL_ONE_COMM")) {
    ncclComm_t dup = nullptr;
    if (g_nccl.CommSplit(g_comm, 0, rank, &dup, nullptr) == ncclSuccess && dup) g_comm_bucket = dup;
  }
  g_world = world;
  g_rank = rank;
  PDN_CUDA(cudaEventCreateWithFlags(&g_ev_compute, cudaEventDisableTiming));
  PDN_CUDA(cudaEventCreateWithFlags(&g_ev_comm, cudaEventDisableTiming));
  return 0;
}

int pdn_nccl_world(int* rank, int* world) {
  *rank = g_rank;
  *world = g_comm ? g_world : 1;
  if (g_comm && g_nccl.CommCount && g_nccl.CommUserRank) {  // what the live communicator itself reports
    PDN_NCCL(g_nccl.CommCount(g_comm_bucket, world));
    PDN_NCCL(g_nccl.CommUserRank(g_comm_bucket, rank));
  }
  return 0;
}

int pdn_allreduce_sum_f32(float* buf, int64_t n) {
  PDN_CHECK(g_comm, "pdn_nccl_init has not been called");
  if (n == 0) return 0;
  // comm stream starts after everything queued on the compute stream so far (the backward pass that filled the bucket)
  PDN_CUDA(cudaEventRecord(g_ev_compute, stream()));
  PDN_CUDA(cudaStreamWaitEvent(comm_stream(), g_ev_compute, 0));
  PDN_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, g_comm_bucket, comm_stream()));
  PDN_CUDA(cudaEventRecord(g_ev_comm, comm_stream()));
  return 0;
}

int pdn_allreduce_wait(void) {
  PDN_CHECK(g_comm, "pdn_nccl_init has not been called");
  PDN_CUDA(cudaStreamWaitEvent(stream(), g_ev_comm, 0));
  return 0;
}

int pdn_allreduce_sum_f32_inline(float* buf, int64_t n) {
  PDN_CHECK(g_comm, "pdn_nccl_init has not been called");
  if (n == 0) return 0;
  PDN_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32, ncclSum, g_comm, stream()));
  return 0;
}

int pdn_nccl_destroy(void) {
  if (g_comm) {
    cudaStreamSynchronize(comm_stream());
    cudaStreamSynchronize(stream());
    if (g_comm_bucket && g_comm_bucket != g_comm) g_nccl.CommDestroy(g_comm_bucket);
    g_nccl.CommDestroy(g_comm);
    g_comm = g_comm_bucket = nullptr;
    g_world = 1;
    g_rank = 0;
  }
  return 0;
}

}  // extern "C"
