// Library plumbing: status strings, device/memory/stream helpers, cached device properties.
#include "common.cuh"

namespace ssdg {
static int g_sm_count[64];
static int g_smem_optin[64];
static int cur_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return d & 63;
}
int sm_count() {
  int d = cur_dev();
  if (g_sm_count[d] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148;
    g_sm_count[d] = v;
  }
  return g_sm_count[d];
}
int max_smem_optin() {
  int d = cur_dev();
  if (g_smem_optin[d] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, d) != cudaSuccess || v <= 0) v = 48 * 1024;
    g_smem_optin[d] = v;
  }
  return g_smem_optin[d];
}
// Optional event bracketing of the dominant kernels (ssdg_profile_enable): one set of events per device, created
// on first use on that device; a measurement aid for single-threaded drivers (bench.py), off by default.
constexpr int kProfSlots = 8;
static bool g_prof = false;
static cudaEvent_t g_prof_ev[64][kProfSlots][2];
static bool g_prof_have[64][kProfSlots];
void prof_begin(int which, cudaStream_t st) {
  if (!g_prof) return;
  const int d = cur_dev();
  if (!g_prof_have[d][which]) {
    cudaEventCreate(&g_prof_ev[d][which][0]);
    cudaEventCreate(&g_prof_ev[d][which][1]);
    g_prof_have[d][which] = true;
  }
  cudaEventRecord(g_prof_ev[d][which][0], st);
}
void prof_end(int which, cudaStream_t st) {
  if (!g_prof) return;
  const int d = cur_dev();
  if (!g_prof_have[d][which]) return;
  cudaEventRecord(g_prof_ev[d][which][1], st);
}
}  // namespace ssdg

extern "C" {

int ssdg_profile_enable(int enable) {
  ssdg::g_prof = enable != 0;
  return SSDG_OK;
}
int ssdg_profile_span_ms(int32_t which, void* ref_event, float* begin_ms, float* end_ms) {
  if (which < 0 || which >= ssdg::kProfSlots || !ref_event || !begin_ms || !end_ms) return SSDG_ERR_ARG;
  const int d = ssdg::cur_dev();
  if (!ssdg::g_prof_have[d][which]) return SSDG_ERR_ARG;
  cudaError_t e = cudaEventSynchronize(ssdg::g_prof_ev[d][which][1]);
  if (e != cudaSuccess) return (int)e;
  e = cudaEventElapsedTime(begin_ms, (cudaEvent_t)ref_event, ssdg::g_prof_ev[d][which][0]);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaEventElapsedTime(end_ms, (cudaEvent_t)ref_event, ssdg::g_prof_ev[d][which][1]);
}
int ssdg_profile_last_ms(int which, float* ms) {
  if (which < 0 || which >= ssdg::kProfSlots || !ms) return SSDG_ERR_ARG;
  const int d = ssdg::cur_dev();
  if (!ssdg::g_prof_have[d][which]) { *ms = 0.f; return SSDG_ERR_ARG; }
  cudaError_t e = cudaEventSynchronize(ssdg::g_prof_ev[d][which][1]);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaEventElapsedTime(ms, ssdg::g_prof_ev[d][which][0], ssdg::g_prof_ev[d][which][1]);
}

const char* ssdg_status_string(int status) {
  switch (status) {
    case SSDG_OK: return "ok";
    case SSDG_ERR_ARG: return "invalid argument";
    case SSDG_ERR_TOO_MANY_GT: return "number of default boxes should greater than the number of targets";
    case SSDG_ERR_THRESH: return "thresh should greater than zero";
    case SSDG_ERR_SHAPE: return "inconsistent shapes";
    case SSDG_ERR_NO_POSITIVE: return "no positive prior in the batch (hard-negative k = 0)";
    case SSDG_ERR_TOPK_RANGE: return "hard-negative k exceeds the number of priors in the batch";
    case SSDG_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
    case SSDG_ERR_ALIGN: return "pointer not 16-byte aligned";
    case SSDG_ERR_LIMIT: return "size beyond an implementation limit";
    case SSDG_ERR_POS_NEG_OVERLAP: return "a positive prior was mined as a hard negative";
    case SSDG_ERR_NO_NCCL: return "libnccl.so.2 could not be loaded (ssdg_comm_* needs NCCL)";
    case SSDG_ERR_LABEL_RANGE: return "class id of a positive prior outside [0, n_classes)";
    case SSDG_ERR_STALE_INDEX: return "the prior index was not built from these priors";
    default: break;
  }
  if (status >= SSDG_ERR_NCCL_BASE) return ssdg::nccl_error_string(status - SSDG_ERR_NCCL_BASE);
  if (status > 0) return cudaGetErrorString((cudaError_t)status);
  return "unknown status";
}

int ssdg_version(void) { return SSDG_VERSION; }

int ssdg_device_count(int* count) {
  if (!count) return SSDG_ERR_ARG;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) { *count = 0; return (int)e; }
  return SSDG_OK;
}
int ssdg_set_device(int device) { return (int)cudaSetDevice(device); }
int ssdg_get_device(int* device) {
  if (!device) return SSDG_ERR_ARG;
  return (int)cudaGetDevice(device);
}
int ssdg_device_alloc(void** dptr, size_t bytes) {
  if (!dptr) return SSDG_ERR_ARG;
  return (int)cudaMalloc(dptr, bytes ? bytes : 1);
}
int ssdg_device_free(void* dptr) { return (int)cudaFree(dptr); }
int ssdg_host_alloc(void** hptr, size_t bytes) {
  if (!hptr) return SSDG_ERR_ARG;
  return (int)cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault);
}
int ssdg_host_free(void* hptr) { return (int)cudaFreeHost(hptr); }
int ssdg_memcpy_h2d(void* dst, const void* src_host, size_t bytes, void* stream) {
  return (int)cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
}
int ssdg_memcpy_d2h(void* dst_host, const void* src, size_t bytes, void* stream) {
  return (int)cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}
int ssdg_memset(void* dst, int value, size_t bytes, void* stream) {
  return (int)cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream);
}
int ssdg_stream_create(void** stream) {
  if (!stream) return SSDG_ERR_ARG;
  cudaStream_t s;
  cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  *stream = (void*)s;
  return (int)e;
}
int ssdg_stream_create_priority(void** stream, int high_priority) {
  if (!stream) return SSDG_ERR_ARG;
  int lo = 0, hi = 0;   // numerically lower = higher priority
  cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (e != cudaSuccess) return (int)e;
  cudaStream_t s;
  // 0: lowest, 1: highest, k >= 2: k-1 levels below the highest (clamped to the lowest)
  int pr = high_priority == 0 ? lo : hi + (high_priority - 1);
  if (pr > lo) pr = lo;
  e = cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, pr);
  *stream = (void*)s;
  return (int)e;
}
int ssdg_stream_destroy(void* stream) { return (int)cudaStreamDestroy((cudaStream_t)stream); }
int ssdg_stream_sync(void* stream) { return (int)cudaStreamSynchronize((cudaStream_t)stream); }
int ssdg_event_create(void** event) {
  if (!event) return SSDG_ERR_ARG;
  cudaEvent_t e;
  cudaError_t r = cudaEventCreate(&e);
  *event = (void*)e;
  return (int)r;
}
int ssdg_event_destroy(void* event) { return (int)cudaEventDestroy((cudaEvent_t)event); }
int ssdg_event_record(void* event, void* stream) { return (int)cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream); }
int ssdg_stream_wait_event(void* stream, void* event) {
  return (int)cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0);
}
int ssdg_event_elapsed_ms(void* start, void* stop, float* ms) {
  if (!ms) return SSDG_ERR_ARG;
  cudaError_t e = cudaEventSynchronize((cudaEvent_t)stop);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop);
}

}  // extern "C"
