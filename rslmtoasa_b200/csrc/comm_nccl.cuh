// comm_nccl.cuh -- the path's one exchange step inside the library: NCCL collectives on device-resident results.
//
// The reference sums per-rank results with MPI_ALLREDUCE(MPI_IN_PLACE, .., MPI_SUM) (bands.f90:270-275, self.f90:887) and
// has an (commented-out) MPI_Allgather of a_b/b2_b (recursion.f90:1788-1799); units are sharded with get_mpi_variables
// (mpi.f90:32-58).  Here the same operations run over NCCL (NVLink 5 / NVSwitch) on the handle's stream, directly on the
// device buffers the kernels wrote, so nothing bounces through host memory before the exchange.
//
// libnccl.so.2 is opened at run time (dlopen) the first time a communicator is requested: librsrec.so itself has no link
// dependency on NCCL, so a single-GPU Fortran host needs nothing but the CUDA runtime.  When the process already carries
// an NCCL (torch's bundled one, or the MPI launcher's), that copy is reused (RTLD_NOLOAD first).
// Only the handful of entry points and enum values the path needs are declared (values as in nccl.h 2.27/2.28).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <string>

#define RS_NCCL_ID_BYTES 128
struct rs_ncclComm;
typedef rs_ncclComm *rs_ncclComm_t;
typedef struct { char internal[RS_NCCL_ID_BYTES]; } rs_ncclUniqueId;
enum { RS_NCCL_SUCCESS = 0, RS_NCCL_SUM = 0, RS_NCCL_INT32 = 2, RS_NCCL_FLOAT64 = 8 };

struct NcclApi {
  void *dl = nullptr;
  int (*GetVersion)(int *) = nullptr;
  int (*GetUniqueId)(rs_ncclUniqueId *) = nullptr;
  int (*CommInitRank)(rs_ncclComm_t *, int, rs_ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(rs_ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, rs_ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, rs_ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, rs_ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  std::string err;
};

static NcclApi *nccl_api() {
  static NcclApi api;
  if (api.dl || !api.err.empty()) return &api;
  const char *names[] = {getenv("RSREC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (int pass = 0; pass < 2 && !api.dl; pass++)
    for (const char *n : names) {
      if (!n) continue;
      api.dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL | (pass == 0 ? RTLD_NOLOAD : 0));
      if (api.dl) break;
    }
  if (!api.dl) { api.err = std::string("cannot open libnccl.so.2 (set RSREC_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return &api; }
#define RS_SYM(field, name)                                             \
  *(void **)(&api.field) = dlsym(api.dl, name);                         \
  if (!api.field) { api.err = std::string("libnccl lacks ") + name; api.dl = nullptr; return &api; }
  RS_SYM(GetVersion, "ncclGetVersion")
  RS_SYM(GetUniqueId, "ncclGetUniqueId")
  RS_SYM(CommInitRank, "ncclCommInitRank")
  RS_SYM(CommDestroy, "ncclCommDestroy")
  RS_SYM(GetErrorString, "ncclGetErrorString")
  RS_SYM(AllReduce, "ncclAllReduce")
  RS_SYM(Broadcast, "ncclBroadcast")
  RS_SYM(AllGather, "ncclAllGather")
  RS_SYM(GroupStart, "ncclGroupStart")
  RS_SYM(GroupEnd, "ncclGroupEnd")
#undef RS_SYM
  return &api;
}
