// K6: data-parallel LoRA-gradient exchange.  Replaces the implicit Lightning-DDP reducer
// (train.py:98-109, modules/utils/fix_ddp.py:5-11) with ONE NCCL all-reduce (average) over the flat
// gradient arena per optimizer step, on the caller's stream (graph-capturable).
//
// libnccl is resolved with dlopen at first use, so libsdt_b200.so itself has no link-time NCCL
// dependency and loads on CPU-only machines (the "not gpu" test tier checks the exported symbols).
#include "sdt_common.cuh"

#include <dlfcn.h>
#include <string.h>

namespace sdt {

// minimal NCCL ABI (stable since 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclBfloat16 = 9 };
enum { ncclSum = 0, ncclAvg = 4 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommFinalize)(ncclComm_t) = nullptr;     // optional (NCCL >= 2.14)
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_world = 0;

static int load_nccl() {
  if (g_nccl.handle) return SDT_OK;
  // torch has normally already mapped its bundled libnccl.so.2; RTLD_NOLOAD picks that copy up first
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (h) break; }
  if (!h) for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  SDT_REQUIRE(h != nullptr, SDT_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SDT_SYM(field, name)                                                         \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                         \
  SDT_REQUIRE(g_nccl.field != nullptr, SDT_ERR_NCCL, "libnccl lacks symbol %s", name)
  SDT_SYM(GetUniqueId, "ncclGetUniqueId");
  SDT_SYM(CommInitRank, "ncclCommInitRank");
  SDT_SYM(AllReduce, "ncclAllReduce");
  SDT_SYM(CommDestroy, "ncclCommDestroy");
  SDT_SYM(GetErrorString, "ncclGetErrorString");
#undef SDT_SYM
  *(void**)(&g_nccl.CommFinalize) = dlsym(h, "ncclCommFinalize");
  g_nccl.handle = h;
  return SDT_OK;
}

#define SDT_NCCL_OK(expr)                                                             \
  do {                                                                                \
    ncclResult_t _r = (expr);                                                         \
    if (_r != 0) {                                                                    \
      set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r));                   \
      return SDT_ERR_NCCL;                                                            \
    }                                                                                 \
  } while (0)

}  // namespace sdt

using namespace sdt;

extern "C" int sdt_comm_unique_id(void* out_128_bytes) {
  SDT_REQUIRE(out_128_bytes, SDT_ERR_ARG, "sdt_comm_unique_id: null pointer");
  int rc = load_nccl();
  if (rc != SDT_OK) return rc;
  ncclUniqueId id;
  SDT_NCCL_OK(g_nccl.GetUniqueId(&id));
  memcpy(out_128_bytes, &id, sizeof(id));
  return SDT_OK;
}

extern "C" int sdt_comm_init(const void* unique_id_128_bytes, int rank, int world) {
  SDT_REQUIRE(unique_id_128_bytes, SDT_ERR_ARG, "sdt_comm_init: null pointer");
  SDT_REQUIRE(world >= 1 && rank >= 0 && rank < world, SDT_ERR_ARG, "sdt_comm_init: bad rank %d / world %d", rank, world);
  SDT_REQUIRE(g_comm == nullptr, SDT_ERR_ARG, "sdt_comm_init: communicator already initialised (one per process)");
  int rc = load_nccl();
  if (rc != SDT_OK) return rc;
  ncclUniqueId id;
  memcpy(&id, unique_id_128_bytes, sizeof(id));
  SDT_NCCL_OK(g_nccl.CommInitRank(&g_comm, world, id, rank));
  g_world = world;
  return SDT_OK;
}

extern "C" int sdt_comm_world(void) { return g_world; }

extern "C" int sdt_allreduce(void* buf, int64_t count, int dtype, void* stream) {
  SDT_REQUIRE(g_comm != nullptr, SDT_ERR_ARG, "sdt_allreduce: call sdt_comm_init first");
  SDT_REQUIRE(buf && count >= 0, SDT_ERR_ARG, "sdt_allreduce: bad buffer");
  SDT_REQUIRE(dtype == SDT_F32 || dtype == SDT_BF16, SDT_ERR_UNSUPPORTED, "sdt_allreduce: unsupported dtype %d", dtype);
  if (count == 0) return SDT_OK;
  // DDP semantics: mean over ranks
  SDT_NCCL_OK(g_nccl.AllReduce(buf, buf, (size_t)count, dtype == SDT_F32 ? ncclFloat32 : ncclBfloat16, ncclAvg, g_comm,
                               (cudaStream_t)stream));
  return SDT_OK;
}

extern "C" int sdt_comm_destroy(void) {
  if (g_comm == nullptr) return SDT_OK;
  // Order matters after the collective has been part of a captured graph: the caller destroys the graph exec first
  // (LatentDiffusionTrainer.release_cuda_graph), then every stream is drained here, the communicator is finalised (flushes
  // its proxy / outstanding work) and only then destroyed.  Destroying with captured work still referenced blocks forever.
  SDT_CUDA_OK(cudaDeviceSynchronize());
  if (g_nccl.CommFinalize != nullptr) SDT_NCCL_OK(g_nccl.CommFinalize(g_comm));
  SDT_NCCL_OK(g_nccl.CommDestroy(g_comm));
  g_comm = nullptr;
  g_world = 0;
  return SDT_OK;
}
