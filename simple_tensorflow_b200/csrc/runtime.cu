// StreamExecutor-level shim of libb200tf.so: streams, events, device/pinned memory, copies,
// plus library-wide helpers (error string, launch counter, driver entry points, NCCL loader).
// Mirrors the members of perftools::gputools::Stream / StreamExecutor that the reference's GPU
// device and kernels use (tensorflow/stream_executor/stream.h:116,189,214,1482-1531,1591;
// stream_executor_pimpl.h:110,191,226-268).
#include <algorithm>
#include <atomic>
#include <map>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <vector>

#include "b200_internal.h"

namespace b200 {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_peer_collectives{0}, g_nccl_collectives{0};
static std::atomic<int> g_matmul_precision{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void note_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
void note_collective(bool peer) {
  (peer ? g_peer_collectives : g_nccl_collectives).fetch_add(1, std::memory_order_relaxed);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return B200_INTERNAL;
  }
  return B200_OK;
}

int require_device(const char* what) {
  static int cached = -1;  // device count; racy init is benign
  if (cached < 0) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
      n = 0;
      cudaGetLastError();
    }
    cached = n;
  }
  if (cached == 0) {
    set_last_error("%s: no CUDA device available (libb200tf has no CPU fallback)", what);
    return B200_INTERNAL;
  }
  return B200_OK;
}

int sm_count() {
  static int cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

const DriverApi& driver() {
  static DriverApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      api.cuTensorMapEncodeTiled = reinterpret_cast<decltype(api.cuTensorMapEncodeTiled)>(fn);
    } else {
      api.cuTensorMapEncodeTiled = nullptr;
      cudaGetLastError();
    }
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      api.cuTensorMapEncodeIm2col = reinterpret_cast<decltype(api.cuTensorMapEncodeIm2col)>(fn);
    } else {
      api.cuTensorMapEncodeIm2col = nullptr;
      cudaGetLastError();
    }
  });
  return api;
}

// ------------------------------------------------------------------ GEMM launch profiling
struct ProfileState {
  std::mutex mu;
  bool enabled = false;
  std::vector<cudaEvent_t> starts, stops;
  std::vector<double> flops;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t pending_start = nullptr;
};
static ProfileState g_prof;
static std::atomic<bool> g_prof_on{false};
bool profile_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
static cudaEvent_t prof_event() {
  if (!g_prof.pool.empty()) {
    cudaEvent_t e = g_prof.pool.back();
    g_prof.pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void profile_gemm_launch_begin(cudaStream_t stream) {
  std::lock_guard<std::mutex> l(g_prof.mu);
  g_prof.pending_start = prof_event();
  cudaEventRecord(g_prof.pending_start, stream);
}
// ---- B200TF_KERNEL_TIMES=1: per-kernel in-stream times (see b200_internal.h)
namespace {
struct KernelTimes {
  std::mutex mu;
  std::vector<cudaEvent_t> pool;
  struct Rec { cudaEvent_t a, b; const void* fn; };
  std::vector<Rec> recs;
  void dump() {
    std::lock_guard<std::mutex> l(mu);
    if (recs.empty()) return;
    cudaDeviceSynchronize();
    struct Tot { double us = 0; long n = 0; };
    std::map<const void*, Tot> tot;
    double all = 0;
    for (auto& r : recs) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
        tot[r.fn].us += ms * 1e3;
        tot[r.fn].n += 1;
        all += ms * 1e3;
      }
    }
    std::vector<std::pair<double, const void*>> order;
    for (auto& kv : tot) order.push_back({kv.second.us, kv.first});
    std::sort(order.begin(), order.end(), [](auto& x, auto& y) { return x.first > y.first; });
    fprintf(stderr, "[b200 kernel times] %zu launches, %.1f us bracketed in total\n", recs.size(), all);
    for (auto& o : order) {
      const char* name = nullptr;
      if (cudaFuncGetName(&name, o.second) != cudaSuccess || !name) name = "?";
      const Tot& t = tot[o.second];
      fprintf(stderr, "  %10.1f us  %6ld x %8.2f us  %5.1f%%  %.110s\n", t.us, t.n, t.us / t.n,
              100.0 * t.us / all, name);
    }
    cudaGetLastError();
    recs.clear();
  }
};
KernelTimes& kernel_times() {
  static KernelTimes* k = [] {
    KernelTimes* p = new KernelTimes();
    atexit([] { kernel_times().dump(); });
    return p;
  }();
  return *k;
}
}  // namespace
bool kernel_times_enabled() {
  static const bool on = getenv("B200TF_KERNEL_TIMES") != nullptr;
  return on;
}
void* kernel_times_begin(cudaStream_t stream) {
  (void)kernel_times();
  cudaEvent_t a = nullptr;
  cudaEventCreate(&a);
  cudaEventRecord(a, stream);
  return a;
}
void kernel_times_end(void* token, cudaStream_t stream, const void* kernel) {
  KernelTimes& k = kernel_times();
  cudaEvent_t b = nullptr;
  cudaEventCreate(&b);
  cudaEventRecord(b, stream);
  std::lock_guard<std::mutex> l(k.mu);
  if (k.recs.size() < (1u << 20)) k.recs.push_back({static_cast<cudaEvent_t>(token), b, kernel});
}

bool pdl_enabled() {
  static const bool on = getenv("B200TF_NO_PDL") == nullptr;
  return on;
}
void profile_gemm_launch_end(cudaStream_t stream, double flops) {
  std::lock_guard<std::mutex> l(g_prof.mu);
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, stream);
  g_prof.starts.push_back(g_prof.pending_start);
  g_prof.stops.push_back(e);
  g_prof.flops.push_back(flops);
  g_prof.pending_start = nullptr;
}

// ------------------------------------------------------------------ NCCL (dlopen'ed)
struct Id128 {
  char bytes[128];
};
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value: 128 bytes*/ Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommUserRank)(void*, int*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
};
static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static bool load_nccl() {
  std::call_once(g_nccl_once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    const char* env = getenv("B200TF_NCCL_LIB");
    void* h = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
    for (int i = 0; !h && names[i]; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    g_nccl.handle = h;
    g_nccl.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank =
        reinterpret_cast<int (*)(void**, int, Id128, int)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.AllReduce =
        reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(
            dlsym(h, "ncclAllReduce"));
    g_nccl.AllGather =
        reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(
            dlsym(h, "ncclAllGather"));
    g_nccl.CommUserRank = reinterpret_cast<int (*)(void*, int*)>(dlsym(h, "ncclCommUserRank"));
    g_nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
    g_nccl.GroupStart = reinterpret_cast<int (*)()>(dlsym(h, "ncclGroupStart"));
    g_nccl.GroupEnd = reinterpret_cast<int (*)()>(dlsym(h, "ncclGroupEnd"));
  });
  if (!g_nccl.handle || !g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) {
    set_last_error("NCCL: libnccl.so.2 could not be loaded (set B200TF_NCCL_LIB)");
    return false;
  }
  return true;
}
static int nccl_rc(int r, const char* what) {
  if (r == 0) return B200_OK;
  set_last_error("%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
  return B200_INTERNAL;
}

}  // namespace b200

using namespace b200;

#define CUDA_RC(expr, what)                                               \
  do {                                                                    \
    cudaError_t _e = (expr);                                              \
    if (_e != cudaSuccess) {                                              \
      set_last_error("%s: %s", what, cudaGetErrorString(_e));             \
      cudaGetLastError();                                                 \
      return _e == cudaErrorMemoryAllocation ? B200_RESOURCE_EXHAUSTED    \
                                             : B200_INTERNAL;             \
    }                                                                     \
  } while (0)

extern "C" {

const char* b200_version(void) { return "b200tf 0.1 (sm_100a)"; }
const char* b200_last_error(void) { return g_err; }

int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
int b200_set_device(int ordinal) {
  CUDA_RC(cudaSetDevice(ordinal), "b200_set_device");
  return B200_OK;
}
uint64_t b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void b200_collective_counts(uint64_t* peer_launches, uint64_t* nccl_calls) {
  if (peer_launches) *peer_launches = g_peer_collectives.load(std::memory_order_relaxed);
  if (nccl_calls) *nccl_calls = g_nccl_collectives.load(std::memory_order_relaxed);
}

int b200_set_matmul_precision(int mode) {
  if (mode < 0 || mode > 1) {
    set_last_error("b200_set_matmul_precision: mode must be 0 or 1 (got %d)", mode);
    return B200_INVALID_ARGUMENT;
  }
  g_matmul_precision.store(mode);
  return B200_OK;
}
int b200_get_matmul_precision(void) { return g_matmul_precision.load(); }

int b200_profile_active(void) {
  return (g_prof_on.load() || b200::kernel_times_enabled()) ? 1 : 0;
}
void b200_note_launches(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void b200_note_collectives(uint64_t peer_launches, uint64_t nccl_calls) {
  g_peer_collectives.fetch_add(peer_launches, std::memory_order_relaxed);
  g_nccl_collectives.fetch_add(nccl_calls, std::memory_order_relaxed);
}

int b200_stream_begin_capture(void* stream) {
  CUDA_RC(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeRelaxed),
          "b200_stream_begin_capture");
  return B200_OK;
}
int b200_stream_end_capture(void* stream, void** graph_exec) {
  *graph_exec = nullptr;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(as_stream(stream), &graph);
  if (e != cudaSuccess || graph == nullptr) {
    set_last_error("b200_stream_end_capture: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return B200_INTERNAL;
  }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    set_last_error("cudaGraphInstantiate: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return B200_INTERNAL;
  }
  *graph_exec = exec;
  return B200_OK;
}
int b200_graph_launch(void* graph_exec, void* stream) {
  CUDA_RC(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), as_stream(stream)),
          "b200_graph_launch");
  return B200_OK;
}
int b200_graph_destroy(void* graph_exec) {
  if (graph_exec) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec));
  cudaGetLastError();
  return B200_OK;
}

int b200_profile_begin(void) {
  std::lock_guard<std::mutex> l(g_prof.mu);
  g_prof.starts.clear();
  g_prof.stops.clear();
  g_prof.flops.clear();
  g_prof.enabled = true;
  g_prof_on.store(true);
  return B200_OK;
}
int b200_profile_end(double* gemm_ms_total, uint64_t* gemm_launches, double* gemm_flops_total) {
  g_prof_on.store(false);
  std::lock_guard<std::mutex> l(g_prof.mu);
  g_prof.enabled = false;
  double ms = 0.0, fl = 0.0;
  for (size_t i = 0; i < g_prof.starts.size(); ++i) {
    cudaEventSynchronize(g_prof.stops[i]);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.starts[i], g_prof.stops[i]) == cudaSuccess) ms += t;
    fl += g_prof.flops[i];
    g_prof.pool.push_back(g_prof.starts[i]);
    g_prof.pool.push_back(g_prof.stops[i]);
  }
  if (gemm_ms_total) *gemm_ms_total = ms;
  if (gemm_launches) *gemm_launches = g_prof.starts.size();
  if (gemm_flops_total) *gemm_flops_total = fl;
  g_prof.starts.clear();
  g_prof.stops.clear();
  g_prof.flops.clear();
  cudaGetLastError();
  return B200_OK;
}

int b200_stream_create(void** stream) {
  cudaStream_t s;
  CUDA_RC(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "b200_stream_create");
  *stream = s;
  return B200_OK;
}
int b200_stream_add_host_callback(void* stream, void (*fn)(void*), void* arg) {
  if (fn == nullptr) {
    set_last_error("b200_stream_add_host_callback: null callback");
    return B200_INVALID_ARGUMENT;
  }
  CUDA_RC(cudaLaunchHostFunc(as_stream(stream), fn, arg), "b200_stream_add_host_callback");
  return B200_OK;
}
int b200_stream_create_with_priority(void** stream, int high_priority) {
  int least = 0, greatest = 0;
  CUDA_RC(cudaDeviceGetStreamPriorityRange(&least, &greatest), "b200_stream_create_with_priority");
  cudaStream_t s;
  CUDA_RC(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? greatest : least),
          "b200_stream_create_with_priority");
  *stream = s;
  return B200_OK;
}
int b200_stream_destroy(void* stream) {
  CUDA_RC(cudaStreamDestroy(as_stream(stream)), "b200_stream_destroy");
  return B200_OK;
}
int b200_stream_synchronize(void* stream) {
  CUDA_RC(cudaStreamSynchronize(as_stream(stream)), "b200_stream_synchronize");
  return B200_OK;
}
int b200_stream_wait_event(void* stream, void* event) {
  CUDA_RC(cudaStreamWaitEvent(as_stream(stream), static_cast<cudaEvent_t>(event), 0),
          "b200_stream_wait_event");
  return B200_OK;
}
int b200_event_create(void** event) {
  cudaEvent_t e;
  CUDA_RC(cudaEventCreate(&e), "b200_event_create");
  *event = e;
  return B200_OK;
}
int b200_event_destroy(void* event) {
  CUDA_RC(cudaEventDestroy(static_cast<cudaEvent_t>(event)), "b200_event_destroy");
  return B200_OK;
}
int b200_event_record(void* event, void* stream) {
  CUDA_RC(cudaEventRecord(static_cast<cudaEvent_t>(event), as_stream(stream)),
          "b200_event_record");
  return B200_OK;
}
int b200_event_synchronize(void* event) {
  CUDA_RC(cudaEventSynchronize(static_cast<cudaEvent_t>(event)), "b200_event_synchronize");
  return B200_OK;
}
int b200_event_query(void* event) {
  cudaError_t e = cudaEventQuery(static_cast<cudaEvent_t>(event));
  if (e == cudaSuccess) return 0;
  if (e == cudaErrorNotReady) {
    cudaGetLastError();
    return 1;
  }
  set_last_error("b200_event_query: %s", cudaGetErrorString(e));
  return -B200_INTERNAL;
}
int b200_event_elapsed_ms(void* start, void* stop, float* ms) {
  CUDA_RC(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)),
          "b200_event_elapsed_ms");
  return B200_OK;
}
int b200_malloc(void** dptr, size_t bytes) {
  *dptr = nullptr;
  if (bytes == 0) return B200_OK;
  CUDA_RC(cudaMalloc(dptr, bytes), "b200_malloc");
  return B200_OK;
}
int b200_free(void* dptr) {
  if (!dptr) return B200_OK;
  CUDA_RC(cudaFree(dptr), "b200_free");
  return B200_OK;
}
int b200_host_malloc(void** hptr, size_t bytes) {
  *hptr = nullptr;
  if (bytes == 0) return B200_OK;
  CUDA_RC(cudaHostAlloc(hptr, bytes, cudaHostAllocDefault), "b200_host_malloc");
  return B200_OK;
}
int b200_host_free(void* hptr) {
  if (!hptr) return B200_OK;
  CUDA_RC(cudaFreeHost(hptr), "b200_host_free");
  return B200_OK;
}
int b200_memcpy_h2d_async(void* dst, const void* src_host, size_t bytes, void* stream) {
  if (bytes == 0) return B200_OK;
  CUDA_RC(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, as_stream(stream)),
          "b200_memcpy_h2d_async");
  return B200_OK;
}
int b200_memcpy_d2h_async(void* dst_host, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return B200_OK;
  CUDA_RC(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)),
          "b200_memcpy_d2h_async");
  return B200_OK;
}
int b200_memcpy_d2d_async(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return B200_OK;
  CUDA_RC(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)),
          "b200_memcpy_d2d_async");
  return B200_OK;
}
int b200_memset_async(void* dst, int byte_value, size_t bytes, void* stream) {
  if (bytes == 0) return B200_OK;
  CUDA_RC(cudaMemsetAsync(dst, byte_value, bytes, as_stream(stream)), "b200_memset_async");
  return B200_OK;
}
int b200_mem_info(size_t* free_bytes, size_t* total_bytes) {
  CUDA_RC(cudaMemGetInfo(free_bytes, total_bytes), "b200_mem_info");
  return B200_OK;
}

// ------------------------------------------------------------------ NCCL
int b200_nccl_unique_id(void* id128_host) {
  if (!load_nccl()) return B200_FAILED_PRECONDITION;
  return nccl_rc(g_nccl.GetUniqueId(id128_host), "ncclGetUniqueId");
}
int b200_nccl_comm_init_rank(void** comm, int nranks, const void* id128_host, int rank) {
  if (!load_nccl()) return B200_FAILED_PRECONDITION;
  Id128 id;
  memcpy(&id, id128_host, sizeof(id));
  return nccl_rc(g_nccl.CommInitRank(comm, nranks, id, rank), "ncclCommInitRank");
}
int b200_nccl_comm_destroy(void* comm) {
  if (!load_nccl()) return B200_FAILED_PRECONDITION;
  return nccl_rc(g_nccl.CommDestroy(comm), "ncclCommDestroy");
}
int b200_nccl_group_start(void) {
  if (!load_nccl() || !g_nccl.GroupStart) return B200_FAILED_PRECONDITION;
  return nccl_rc(g_nccl.GroupStart(), "ncclGroupStart");
}
int b200_nccl_group_end(void) {
  if (!load_nccl() || !g_nccl.GroupEnd) return B200_FAILED_PRECONDITION;
  return nccl_rc(g_nccl.GroupEnd(), "ncclGroupEnd");
}
int b200_nccl_all_reduce(int dtype, const void* sendbuf, void* recvbuf, int64_t count, int average,
                         void* comm, void* stream) {
  if (!load_nccl()) return B200_FAILED_PRECONDITION;
  note_collective(false);
  int nccl_type;
  if (dtype == B200_DT_FLOAT)
    nccl_type = 7;
  else if (dtype == B200_DT_BFLOAT16)
    nccl_type = 9;
  else {
    set_last_error("b200_nccl_all_reduce: unsupported dtype %d", dtype);
    return B200_UNIMPLEMENTED;
  }
  if (count == 0) return B200_OK;
  return nccl_rc(g_nccl.AllReduce(sendbuf, recvbuf, (size_t)count, nccl_type,
                                  average ? /*ncclAvg*/ 4 : /*ncclSum*/ 0, comm, as_stream(stream)),
                 "ncclAllReduce");
}
int b200_nccl_comm_user_rank(void* comm, int* rank) {
  if (!load_nccl() || !g_nccl.CommUserRank) return B200_FAILED_PRECONDITION;
  return nccl_rc(g_nccl.CommUserRank(comm, rank), "ncclCommUserRank");
}
int b200_nccl_all_gather_bytes(const void* sendbuf, void* recvbuf, int64_t bytes_per_rank,
                               void* comm, void* stream) {
  if (!load_nccl() || !g_nccl.AllGather) return B200_FAILED_PRECONDITION;
  return nccl_rc(g_nccl.AllGather(sendbuf, recvbuf, (size_t)bytes_per_rank, /*ncclInt8*/ 0, comm,
                                  as_stream(stream)),
                 "ncclAllGather");
}
int b200_nccl_all_reduce_sum(int dtype, const void* sendbuf, void* recvbuf, int64_t count,
                             void* comm, void* stream) {
  if (!load_nccl()) return B200_FAILED_PRECONDITION;
  note_collective(false);
  int nccl_type;
  if (dtype == B200_DT_FLOAT)
    nccl_type = 7;  // ncclFloat32
  else if (dtype == B200_DT_BFLOAT16)
    nccl_type = 9;  // ncclBfloat16
  else {
    set_last_error("b200_nccl_all_reduce_sum: unsupported dtype %d", dtype);
    return B200_UNIMPLEMENTED;
  }
  if (count == 0) return B200_OK;
  return nccl_rc(g_nccl.AllReduce(sendbuf, recvbuf, (size_t)count, nccl_type, /*ncclSum*/ 0, comm,
                                  as_stream(stream)),
                 "ncclAllReduce");
}

}  // extern "C"
