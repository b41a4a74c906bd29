// Data-parallel exchange of the hot path in the C ABI: ssdg_comm_* over NCCL.
//
// The path shards by image, so the only exchange is the loss's: the 7 additive words of the result block
// (per-shard mining, models/ssd_model.py:235-256 semantics) or the radix-select histograms of the exact
// batch-global mining (:368-372).  Payloads are <= 8 KB and latency-bound; a plain in-place ncclAllReduce on the
// caller's stream is the right tool (a fused compute+collective kernel has nothing to overlap here).
//
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy a host framework already loaded, else the system one),
// so libssdgeom.so itself has no NCCL dependency and single-GPU callers never touch it.  Types come from <nccl.h>.
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <mutex>
#include "common.cuh"

namespace ssdg {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static void nccl_load() {
  const char* names[] = {getenv("SSDGEOM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) return;
  auto sym = [&](const char* s) { return dlsym(g_nccl.handle, s); };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.CommCount = (decltype(g_nccl.CommCount))sym("ncclCommCount");
  g_nccl.GetVersion = (decltype(g_nccl.GetVersion))sym("ncclGetVersion");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy && g_nccl.GroupStart &&
              g_nccl.GroupEnd;
}

static const NcclApi* nccl() {
  std::call_once(g_nccl_once, nccl_load);
  return g_nccl.ok ? &g_nccl : nullptr;
}

const char* nccl_error_string(int code) {
  const NcclApi* api = nccl();
  if (!api || !api->GetErrorString) return "NCCL error";
  return api->GetErrorString((ncclResult_t)code);
}

struct Comm {
  ncclComm_t comm;
  int world, rank, device;
};

static int nccl_status(ncclResult_t r) { return r == ncclSuccess ? SSDG_OK : SSDG_ERR_NCCL_BASE + (int)r; }

}  // namespace ssdg

using namespace ssdg;

extern "C" {

int ssdg_comm_available(int* nccl_version) {
  const NcclApi* api = nccl();
  if (nccl_version) {
    *nccl_version = 0;
    if (api && api->GetVersion) api->GetVersion(nccl_version);
  }
  return api ? SSDG_OK : SSDG_ERR_NO_NCCL;
}

int ssdg_comm_unique_id(void* id_out) {
  if (!id_out) return SSDG_ERR_ARG;
  const NcclApi* api = nccl();
  if (!api) return SSDG_ERR_NO_NCCL;
  static_assert(sizeof(ncclUniqueId) == SSDG_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  const ncclResult_t r = api->GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_status(r);
  std::memcpy(id_out, &id, sizeof(id));
  return SSDG_OK;
}

int ssdg_comm_init_rank(void** comm_out, const void* id, int32_t world, int32_t rank) {
  if (!comm_out || !id || world <= 0 || rank < 0 || rank >= world) return SSDG_ERR_ARG;
  const NcclApi* api = nccl();
  if (!api) return SSDG_ERR_NO_NCCL;
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  Comm* c = new Comm();
  c->world = world; c->rank = rank;
  SSDG_CUDA_TRY(cudaGetDevice(&c->device));
  const ncclResult_t r = api->CommInitRank(&c->comm, world, uid, rank);
  if (r != ncclSuccess) { delete c; return nccl_status(r); }
  *comm_out = c;
  return SSDG_OK;
}

int ssdg_comm_world(void* comm, int32_t* world, int32_t* rank) {
  if (!comm) return SSDG_ERR_ARG;
  const Comm* c = (const Comm*)comm;
  if (world) *world = c->world;
  if (rank) *rank = c->rank;
  return SSDG_OK;
}

static int dtype_of(int32_t dtype, ncclDataType_t* out) {
  switch (dtype) {
    case SSDG_F32: *out = ncclFloat32; return SSDG_OK;
    case SSDG_F64: *out = ncclFloat64; return SSDG_OK;
    case SSDG_I32: *out = ncclInt32; return SSDG_OK;
    case SSDG_I64: *out = ncclInt64; return SSDG_OK;
    default: return SSDG_ERR_ARG;
  }
}

int ssdg_comm_allreduce_sum(void* comm, void* buf, int64_t count, int32_t dtype, void* stream) {
  if (!comm || !buf || count <= 0) return SSDG_ERR_ARG;
  const NcclApi* api = nccl();
  if (!api) return SSDG_ERR_NO_NCCL;
  ncclDataType_t dt;
  if (dtype_of(dtype, &dt) != SSDG_OK) return SSDG_ERR_ARG;
  Comm* c = (Comm*)comm;
  return nccl_status(api->AllReduce(buf, buf, (size_t)count, dt, ncclSum, c->comm, (cudaStream_t)stream));
}

// Several in-place sums as ONE NCCL group (one launch, one latency): the stage exchanges of the cross-shard mining
// sum a histogram and a count together.
int ssdg_comm_allreduce_sum_multi(void* comm, int32_t n, void* const* bufs, const int64_t* counts, const int32_t* dtypes,
                                  void* stream) {
  if (!comm || n <= 0 || !bufs || !counts || !dtypes) return SSDG_ERR_ARG;
  const NcclApi* api = nccl();
  if (!api) return SSDG_ERR_NO_NCCL;
  Comm* c = (Comm*)comm;
  for (int i = 0; i < n; ++i) {
    ncclDataType_t dt;
    if (!bufs[i] || counts[i] <= 0 || dtype_of(dtypes[i], &dt) != SSDG_OK) return SSDG_ERR_ARG;
  }
  ncclResult_t r = api->GroupStart();
  if (r != ncclSuccess) return nccl_status(r);
  ncclResult_t first = ncclSuccess;
  for (int i = 0; i < n; ++i) {
    ncclDataType_t dt;
    dtype_of(dtypes[i], &dt);
    r = api->AllReduce(bufs[i], bufs[i], (size_t)counts[i], dt, ncclSum, c->comm, (cudaStream_t)stream);
    if (r != ncclSuccess && first == ncclSuccess) first = r;
  }
  r = api->GroupEnd();
  return nccl_status(first != ncclSuccess ? first : r);
}

int ssdg_comm_destroy(void* comm) {
  if (!comm) return SSDG_OK;
  const NcclApi* api = nccl();
  Comm* c = (Comm*)comm;
  int rc = SSDG_OK;
  if (api) rc = nccl_status(api->CommDestroy(c->comm));
  delete c;
  return rc;
}

}  // extern "C"
