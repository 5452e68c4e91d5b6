// gtn_comm.cu -- thin C-ABI wrappers of the collectives the sharded coarse-graining step uses (SURVEY.md section 8(b):
// gtn_comm_init / gtn_allgather / gtn_allreduce), NCCL over NVLink 5 / NVSwitch.
//
// The reference has no distributed mode; these replace nothing upstream.  They exist so that a
// host that is not Python / torch.distributed can drive the sharded step (grassmanntn_b200/sharded.py: all-reduce of the
// l x p sketch panels and l x l Gram matrices, all-gather of the isometries and of the projected-matrix rows,
// broadcast of the owners' small SVDs) with the same C ABI as the kernels.  NCCL is bound at RUN time (dlopen of
// libnccl.so.2, preferring a copy the process has already loaded, e.g. the one bundled with torch), so the library
// builds, loads and passes its CPU checks on a box without NCCL or GPUs.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include "../../include/gtn_b200.h"

namespace {

// the slice of the NCCL ABI used here (stable since NCCL 2.x)
typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm_t;
enum { kNcclFloat64 = 8 };
enum { kNcclSum = 0, kNcclProd = 1, kNcclMax = 2, kNcclMin = 3 };

struct Api {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_unique_id*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  bool ok = false;
};

Api& api() {
  static Api a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy already in the process, if any
  if (!a.handle) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!a.handle) return a;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
  a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
  a.Broadcast = (decltype(a.Broadcast))dlsym(a.handle, "ncclBroadcast");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.Broadcast;
  return a;
}

inline int64_t doubles_of(int64_t count, int dtype) { return dtype == GTN_C128 ? 2 * count : count; }
inline int nccl_rc(int rc) { return rc == 0 ? GTN_OK : 1000 + rc; }      // NCCL errors are reported as 1000 + ncclResult_t

// Device-side barrier between the GPUs of a box over peer-mapped memory: every rank writes `epoch` into slot `rank` of
// EVERY rank's flag array and waits until all slots of its own array have reached `epoch`.  System-scope fences on both
// sides order the peer stores of the kernels launched before the barrier (the GEMM epilogue's partial panels) before
// the flag, and the consumer kernels after it.  One warp, ~5 us over NVSwitch.
struct FlagPeers { unsigned long long* p[GTN_MAX_PEERS]; };

__global__ void peer_barrier_kernel(FlagPeers flags, int rank, int npeers, unsigned long long epoch) {
  const int t = threadIdx.x;
  if (t < npeers) {
    __threadfence_system();
    volatile unsigned long long* dst = flags.p[t] + rank;
    *dst = epoch;
    volatile unsigned long long* src = flags.p[rank] + t;
    const long long t0 = clock64();
    while (*src < epoch) {
      if (clock64() - t0 > 20000000000ll) break;       // ~10 s: a peer died; do not hang the GPU (the results are
    }                                                  // then wrong and the step's certificate / norm checks fail)
    __threadfence_system();
  }
}

}  // namespace

extern "C" int gtn_peer_barrier(void* const* flag_peers, int rank, int npeers, uint64_t epoch, void* stream) {
  if (!flag_peers || npeers < 1 || npeers > GTN_MAX_PEERS || rank < 0 || rank >= npeers) return GTN_ERR_BAD_ARG;
  FlagPeers f;
  for (int i = 0; i < GTN_MAX_PEERS; ++i) f.p[i] = i < npeers ? (unsigned long long*)flag_peers[i] : nullptr;
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, npeers, (unsigned long long)epoch);
  return (int)cudaGetLastError();
}

extern "C" int gtn_comm_available(void) { return api().ok ? 1 : 0; }

extern "C" int gtn_comm_unique_id(void* id_out) {
  if (!id_out) return GTN_ERR_BAD_ARG;
  if (!api().ok) return GTN_ERR_UNSUPPORTED;
  nccl_unique_id id;
  const int rc = api().GetUniqueId(&id);
  if (rc == 0) memcpy(id_out, &id, sizeof(id));
  return nccl_rc(rc);
}

extern "C" int gtn_comm_init(const void* id, int rank, int world, void** comm_out) {
  if (!id || !comm_out || world < 1 || rank < 0 || rank >= world) return GTN_ERR_BAD_ARG;
  if (!api().ok) return GTN_ERR_UNSUPPORTED;
  nccl_unique_id uid;
  memcpy(&uid, id, sizeof(uid));
  nccl_comm_t c = nullptr;
  const int rc = api().CommInitRank(&c, world, uid, rank);
  *comm_out = c;
  return nccl_rc(rc);
}

extern "C" int gtn_comm_destroy(void* comm) {
  if (!comm) return GTN_OK;
  if (!api().ok) return GTN_ERR_UNSUPPORTED;
  return nccl_rc(api().CommDestroy((nccl_comm_t)comm));
}

extern "C" int gtn_allreduce(void* comm, void* buf, int64_t count, int dtype, int op, void* stream) {
  if (!comm || (dtype != GTN_C128 && dtype != GTN_F64) || op < 0 || op > 2) return GTN_ERR_BAD_ARG;
  if (dtype == GTN_C128 && op != 0) return GTN_ERR_BAD_ARG;              // max / min are not defined on complex numbers
  if (!api().ok) return GTN_ERR_UNSUPPORTED;
  if (count <= 0) return GTN_OK;
  const int nop = op == 0 ? kNcclSum : (op == 1 ? kNcclMax : kNcclMin);
  return nccl_rc(api().AllReduce(buf, buf, (size_t)doubles_of(count, dtype), kNcclFloat64, nop, (nccl_comm_t)comm,
                                 (cudaStream_t)stream));
}

extern "C" int gtn_allgather(void* comm, const void* send, void* recv, int64_t count, int dtype, void* stream) {
  if (!comm || (dtype != GTN_C128 && dtype != GTN_F64)) return GTN_ERR_BAD_ARG;
  if (!api().ok) return GTN_ERR_UNSUPPORTED;
  if (count <= 0) return GTN_OK;
  return nccl_rc(api().AllGather(send, recv, (size_t)doubles_of(count, dtype), kNcclFloat64, (nccl_comm_t)comm,
                                 (cudaStream_t)stream));
}

extern "C" int gtn_broadcast(void* comm, void* buf, int64_t count, int dtype, int root, void* stream) {
  if (!comm || (dtype != GTN_C128 && dtype != GTN_F64) || root < 0) return GTN_ERR_BAD_ARG;
  if (!api().ok) return GTN_ERR_UNSUPPORTED;
  if (count <= 0) return GTN_OK;
  return nccl_rc(api().Broadcast(buf, buf, (size_t)doubles_of(count, dtype), kNcclFloat64, root, (nccl_comm_t)comm,
                                 (cudaStream_t)stream));
}
