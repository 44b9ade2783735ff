// C-ABI collectives of the data-parallel hot path for hosts that are not Python: a thin layer
// over an ncclComm_t, replacing the Horovod calls of the reference
//   hvd.allgather            alpro_models.py:110-111, 764-765   -> alpro_comm_allgather      (rank-order concatenation)
//   its gradient             (Horovod 0.19.4: allreduce + narrow) -> alpro_comm_reduce_scatter (sum, local slice)
//   hvd.DistributedOptimizer run_video_retrieval.py:320-323,444  -> alpro_comm_allreduce      (sum or average, in place)
// NCCL is bound at RUN time (dlopen of libnccl.so.2): inside a PyTorch process that resolves to the NCCL torch already
// loaded (one NCCL per process), elsewhere to the system library; libalpro_b200.so itself has no NCCL link dependency,
// so single-GPU users never need it. Only the handful of entry points below are used; their ABI is stable across
// NCCL 2.x.
#include <dlfcn.h>

#include <mutex>

#include "common.h"

namespace alpro {
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0 };
enum { kNcclSum = 0, kNcclAvg = 4 };
enum { kNcclFloat16 = 6, kNcclFloat32 = 7, kNcclBfloat16 = 9 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) return;
#define ALPRO_SYM(field, sym) *reinterpret_cast<void**>(&a.field) = dlsym(a.lib, sym)
    ALPRO_SYM(GetUniqueId, "ncclGetUniqueId");
    ALPRO_SYM(CommInitRank, "ncclCommInitRank");
    ALPRO_SYM(CommDestroy, "ncclCommDestroy");
    ALPRO_SYM(AllGather, "ncclAllGather");
    ALPRO_SYM(ReduceScatter, "ncclReduceScatter");
    ALPRO_SYM(AllReduce, "ncclAllReduce");
    ALPRO_SYM(GetErrorString, "ncclGetErrorString");
#undef ALPRO_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.ReduceScatter && a.AllReduce;
  });
  return a;
}

struct Comm {
  ncclComm_t nccl;
  int world, rank;
};

int nccl_dtype(int dtype) {   // ALPRO_DT_*: 0 = f32, 1 = f16, 2 = bf16 (the `kind` codes of the elementwise entry points)
  return dtype == 0 ? kNcclFloat32 : dtype == 1 ? kNcclFloat16 : dtype == 2 ? kNcclBfloat16 : -1;
}

int fail(const char* what, int rc) {
  NcclApi& a = api();
  set_last_error("%s: NCCL error %d (%s)", what, rc, a.GetErrorString ? a.GetErrorString(rc) : "?");
  return ALPRO_EINVAL;
}

}  // namespace
}  // namespace alpro

using namespace alpro;

#define ALPRO_NCCL_READY(name)                                                                    \
  do {                                                                                            \
    if (!api().ok) {                                                                              \
      set_last_error("%s: libnccl.so.2 could not be loaded (%s)", name, dlerror() ? dlerror() : "missing symbols"); \
      return ALPRO_ENOTSUP;                                                                       \
    }                                                                                             \
  } while (0)

extern "C" int alpro_comm_unique_id(void* id128) {
  ALPRO_REQUIRE(id128, "alpro_comm_unique_id: null buffer");
  ALPRO_NCCL_READY("alpro_comm_unique_id");
  ncclUniqueId id;
  const int rc = api().GetUniqueId(&id);
  if (rc != kNcclSuccess) return fail("alpro_comm_unique_id", rc);
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int alpro_comm_init(void** comm_out, int world, int rank, const void* id128) {
  ALPRO_REQUIRE(comm_out && id128 && world > 0 && rank >= 0 && rank < world, "alpro_comm_init: bad args");
  ALPRO_NCCL_READY("alpro_comm_init");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  Comm* c = new Comm{nullptr, world, rank};
  const int rc = api().CommInitRank(&c->nccl, world, id, rank);   // binds to the CURRENT CUDA device of the caller
  if (rc != kNcclSuccess) {
    delete c;
    return fail("alpro_comm_init", rc);
  }
  *comm_out = c;
  return 0;
}

extern "C" int alpro_comm_destroy(void* comm) {
  if (!comm) return 0;
  Comm* c = static_cast<Comm*>(comm);
  const int rc = api().ok ? api().CommDestroy(c->nccl) : 0;
  delete c;
  return rc == kNcclSuccess ? 0 : fail("alpro_comm_destroy", rc);
}

extern "C" int alpro_comm_allgather(void* comm, const void* send, void* recv, int64_t count_per_rank, int dtype,
                                    void* stream) {
  ALPRO_REQUIRE(comm && send && recv && count_per_rank > 0 && nccl_dtype(dtype) >= 0, "alpro_comm_allgather: bad args");
  Comm* c = static_cast<Comm*>(comm);
  const int rc = api().AllGather(send, recv, static_cast<size_t>(count_per_rank), nccl_dtype(dtype), c->nccl,
                                 static_cast<cudaStream_t>(stream));
  return rc == kNcclSuccess ? 0 : fail("alpro_comm_allgather", rc);
}

extern "C" int alpro_comm_reduce_scatter(void* comm, const void* send, void* recv, int64_t count_per_rank, int dtype,
                                         void* stream) {
  ALPRO_REQUIRE(comm && send && recv && count_per_rank > 0 && nccl_dtype(dtype) >= 0,
                "alpro_comm_reduce_scatter: bad args");
  Comm* c = static_cast<Comm*>(comm);
  const int rc = api().ReduceScatter(send, recv, static_cast<size_t>(count_per_rank), nccl_dtype(dtype), kNcclSum, c->nccl,
                                     static_cast<cudaStream_t>(stream));
  return rc == kNcclSuccess ? 0 : fail("alpro_comm_reduce_scatter", rc);
}

extern "C" int alpro_comm_allreduce(void* comm, void* buf, int64_t count, int dtype, int average, void* stream) {
  ALPRO_REQUIRE(comm && buf && count > 0 && nccl_dtype(dtype) >= 0, "alpro_comm_allreduce: bad args");
  Comm* c = static_cast<Comm*>(comm);
  const int rc = api().AllReduce(buf, buf, static_cast<size_t>(count), nccl_dtype(dtype), average ? kNcclAvg : kNcclSum,
                                 c->nccl, static_cast<cudaStream_t>(stream));
  return rc == kNcclSuccess ? 0 : fail("alpro_comm_allreduce", rc);
}

extern "C" int alpro_comm_rank(void* comm) { return comm ? static_cast<Comm*>(comm)->rank : -1; }
extern "C" int alpro_comm_world(void* comm) { return comm ? static_cast<Comm*>(comm)->world : -1; }
