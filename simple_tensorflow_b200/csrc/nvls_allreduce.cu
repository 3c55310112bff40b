// All-reduce through NVSwitch multicast memory (NVLS): the switch adds the replicas' values while
// they cross it (multimem.ld_reduce) and fans a store out to every replica (multimem.st), so a
// rank moves 2 x (1 / nranks) of the buffer over its links instead of 2 x (nranks - 1) / nranks
// for the peer-load two-shot of peer_allreduce.cu.  One process per GPU; the multicast object is
// shared as a POSIX file descriptor over a Unix-domain socket (SCM_RIGHTS), every rank binds its own
// VMM allocation to it and maps both views: unicast (what its kernels write gradients into) and
// multicast (what this kernel reduces through).
//
//   rank r owns slice r of the arena window:
//     barrier (multimem.red on a flag word: every replica of the flag counts every rank)
//     x = multimem.ld_reduce.add(slice r)   -- summed in the switch, fp32
//     multimem.st(slice r, x * scale)       -- lands in all nranks replicas
//     barrier
//
// Only the owner reduces a slice, so all replicas hold identical bits.  Reference role: the
// gradient exchange the reference's tower pattern does with _Send/_Recv + AddN
// (core/kernels/aggregate_ops.cc:153-176; third_party/nccl.BUILD has no call sites).
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda_bf16.h>

#include "b200_internal.h"

namespace b200 {

namespace {

constexpr size_t kNvlsHeaderBytes = 64 << 10;  // flag words live in the first 64 KB
constexpr size_t kNvlsCallsOffset = 32 << 10;  // per-CTA barrier counters (this rank's view only)
constexpr int kNvlsMaxCtas = 256;
constexpr int kNvlsThreads = 256;

struct VmmApi {
  bool ok = false;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle,
                               size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*,
                                      CUmulticastGranularity_flags) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*,
                        unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle,
                                         CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*,
                                           CUmemAllocationHandleType) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle,
                     unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*,
                                          CUmemAllocationGranularity_flags) = nullptr;
};

template <typename F>
static bool entry(const char* name, F* fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    return false;
  }
  *fn = reinterpret_cast<F>(p);
  return true;
}

static const VmmApi& vmm() {
  static VmmApi api = [] {
    VmmApi a;
    a.ok = entry("cuDeviceGetAttribute", &a.DeviceGetAttribute) &&
           entry("cuMulticastCreate", &a.MulticastCreate) &&
           entry("cuMulticastAddDevice", &a.MulticastAddDevice) &&
           entry("cuMulticastBindMem", &a.MulticastBindMem) &&
           entry("cuMulticastGetGranularity", &a.MulticastGetGranularity) &&
           entry("cuMulticastUnbind", &a.MulticastUnbind) &&
           entry("cuMemCreate", &a.MemCreate) && entry("cuMemRelease", &a.MemRelease) &&
           entry("cuMemExportToShareableHandle", &a.MemExportToShareableHandle) &&
           entry("cuMemImportFromShareableHandle", &a.MemImportFromShareableHandle) &&
           entry("cuMemAddressReserve", &a.MemAddressReserve) &&
           entry("cuMemAddressFree", &a.MemAddressFree) && entry("cuMemMap", &a.MemMap) &&
           entry("cuMemUnmap", &a.MemUnmap) && entry("cuMemSetAccess", &a.MemSetAccess) &&
           entry("cuMemGetAllocationGranularity", &a.MemGetAllocationGranularity);
    return a;
  }();
  return api;
}

// ---- fd passing over an abstract-namespace Unix socket
static int send_fd(int sock, int fd, int status) {
  struct msghdr msg;
  memset(&msg, 0, sizeof(msg));
  struct iovec iov;
  iov.iov_base = &status;
  iov.iov_len = sizeof(status);
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  if (fd >= 0) {
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    struct cmsghdr* c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
  }
  return sendmsg(sock, &msg, 0) == (ssize_t)sizeof(status) ? 0 : -1;
}
static int recv_fd(int sock, int* status) {
  struct msghdr msg;
  memset(&msg, 0, sizeof(msg));
  struct iovec iov;
  iov.iov_base = status;
  iov.iov_len = sizeof(*status);
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  if (recvmsg(sock, &msg, 0) != (ssize_t)sizeof(*status)) return -1;
  for (struct cmsghdr* c = CMSG_FIRSTHDR(&msg); c != nullptr; c = CMSG_NXTHDR(&msg, c))
    if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
      int fd = -1;
      memcpy(&fd, CMSG_DATA(c), sizeof(int));
      return fd;
    }
  return -1;
}
static socklen_t abstract_addr(const char* name, struct sockaddr_un* addr) {
  memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  const size_t n = strlen(name);
  memcpy(addr->sun_path + 1, name, n);  // leading NUL: abstract namespace, nothing on disk
  return (socklen_t)(offsetof(struct sockaddr_un, sun_path) + 1 + n);
}

__device__ __forceinline__ void multimem_red_add_release(uint32_t* mc_addr, uint32_t v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc_addr, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// CTA b of every rank meets CTA b of the others: one multimem.red bumps flag b in EVERY replica, so
// a rank's own copy reaches `target` (= barriers so far x nranks) once all ranks have arrived.
__device__ __forceinline__ void nvls_barrier(uint32_t* mc_flags, const uint32_t* uc_flags,
                                             uint32_t target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    multimem_red_add_release(mc_flags + blockIdx.x, 1u);
    while (ld_acquire_sys_u32(uc_flags + blockIdx.x) < target) {
    }
  }
  __syncthreads();
}

// 16 bytes of bf16 (8 values) reduced in the switch with fp32 accumulation
__device__ __forceinline__ uint4 multimem_ld_reduce_add_bf16(const void* mc_addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_b32x4(void* mc_addr, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr),
               "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
               "f"(__uint_as_float(v.w))
               : "memory");
}
__device__ __forceinline__ uint32_t scale_bf16x2(uint32_t w, float scale) {
  const float lo = __uint_as_float(w << 16) * scale, hi = __uint_as_float(w & 0xFFFF0000u) * scale;
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// data: 16-byte vectors (4 x fp32 or 8 x bf16) of the window, in place
template <bool kBf16>
__global__ void __launch_bounds__(kNvlsThreads, 4)
nvls_all_reduce_kernel(char* __restrict__ mc_data, uint32_t* mc_flags, const uint32_t* uc_flags,
                       uint32_t* uc_calls, long long nvec, long long slice_vecs, int rank,
                       int nranks, float scale) {
  pdl_prologue();
  // Barrier targets come from a per-CTA counter in this rank's own header (unicast view only): a
  // kernel argument would be frozen into a captured CUDA graph.  Flag word b only ever grows by
  // nranks per barrier CTA b takes part in, on every rank alike.
  __shared__ uint32_t calls_sm;
  if (threadIdx.x == 0) {
    calls_sm = uc_calls[blockIdx.x];
    uc_calls[blockIdx.x] = calls_sm + 2;
  }
  __syncthreads();
  const uint32_t target_begin = (calls_sm + 1) * (uint32_t)nranks;
  const uint32_t target_end = (calls_sm + 2) * (uint32_t)nranks;
  nvls_barrier(mc_flags, uc_flags, target_begin);  // every rank's producers are done
  const long long lo = (long long)rank * slice_vecs;
  long long hi = lo + slice_vecs;
  if (hi > nvec) hi = nvec;
  const long long stride = (long long)gridDim.x * kNvlsThreads;
  constexpr int U = 8;  // reductions in flight per thread (a switch round trip each)
  for (long long v0 = lo + (long long)blockIdx.x * kNvlsThreads + threadIdx.x; v0 < hi;
       v0 += stride * U) {
    uint4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * stride;
      if (v < hi) {
        if (kBf16) {
          x[u] = multimem_ld_reduce_add_bf16(mc_data + 16 * v);
        } else {
          const float4 f = multimem_ld_reduce_add(reinterpret_cast<const float*>(mc_data + 16 * v));
          x[u] = make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z),
                            __float_as_uint(f.w));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * stride;
      if (v < hi) {
        if (kBf16) {
          x[u].x = scale_bf16x2(x[u].x, scale);
          x[u].y = scale_bf16x2(x[u].y, scale);
          x[u].z = scale_bf16x2(x[u].z, scale);
          x[u].w = scale_bf16x2(x[u].w, scale);
        } else {
          x[u].x = __float_as_uint(__uint_as_float(x[u].x) * scale);
          x[u].y = __float_as_uint(__uint_as_float(x[u].y) * scale);
          x[u].z = __float_as_uint(__uint_as_float(x[u].z) * scale);
          x[u].w = __float_as_uint(__uint_as_float(x[u].w) * scale);
        }
        multimem_st_b32x4(mc_data + 16 * v, x[u]);
      }
    }
  }
  __threadfence_system();
  nvls_barrier(mc_flags, uc_flags, target_end);  // every slice has landed in every replica
}

}  // namespace

struct NvlsArena {
  int device = 0, rank = 0, nranks = 0;
  size_t size = 0, data_bytes = 0;
  CUdeviceptr uc_va = 0, mc_va = 0;
  CUmemGenericAllocationHandle mem = 0, mc = 0;
  bool mem_valid = false, mc_valid = false, uc_mapped = false, mc_mapped = false, bound = false;
  uint32_t barriers = 0;  // barriers issued so far (all ranks issue the same call sequence)
};

void nvls_arena_destroy(NvlsArena* a) {
  if (!a) return;
  const VmmApi& v = vmm();
  cudaSetDevice(a->device);
  cudaDeviceSynchronize();
  if (a->mc_mapped) v.MemUnmap(a->mc_va, a->size);
  if (a->mc_va) v.MemAddressFree(a->mc_va, a->size);
  if (a->uc_mapped) v.MemUnmap(a->uc_va, a->size);
  if (a->uc_va) v.MemAddressFree(a->uc_va, a->size);
  if (a->bound) v.MulticastUnbind(a->mc, a->device, 0, a->size);
  if (a->mem_valid) v.MemRelease(a->mem);
  if (a->mc_valid) v.MemRelease(a->mc);
  cudaGetLastError();
  delete a;
}

// all-reduce (sum) of one float over the communicator: the votes that keep every rank's decision
// identical.  Returns the sum, or -1 on failure.
static int vote_sum(void* nccl_comm, int mine) {
  float* d = nullptr;
  if (cudaMalloc(&d, sizeof(float)) != cudaSuccess) {
    cudaGetLastError();
    d = nullptr;
  }
  const float v = mine ? 1.f : 0.f;
  float sum = -1.f;
  if (d) cudaMemcpy(d, &v, sizeof(float), cudaMemcpyHostToDevice);
  // every rank must enter the collective, even one whose cudaMalloc failed (it then hangs the
  // others no longer than NCCL's own error handling would): treat that as fatal for NVLS only
  if (d && b200_nccl_all_reduce_sum(B200_DT_FLOAT, d, d, 1, nccl_comm, nullptr) == B200_OK &&
      cudaDeviceSynchronize() == cudaSuccess)
    cudaMemcpy(&sum, d, sizeof(float), cudaMemcpyDeviceToHost);
  if (d) cudaFree(d);
  cudaGetLastError();
  return sum < 0 ? -1 : (int)(sum + 0.5f);
}

// Collective over `nccl_comm`.  B200_OK with *out set, or B200_UNAVAILABLE on EVERY rank.
int nvls_arena_create(void* nccl_comm, int rank, int nranks, size_t data_bytes, NvlsArena** out) {
  *out = nullptr;
  const VmmApi& v = vmm();
  NvlsArena* a = new NvlsArena();
  a->rank = rank;
  a->nranks = nranks;
  cudaGetDevice(&a->device);
  a->data_bytes = (data_bytes + 255) / 256 * 256;
  CUdevice dev = a->device;  // CUdevice ordinals follow the runtime's for the primary contexts
  // ---- phase 1: can this rank do it at all?
  int ok = v.ok ? 1 : 0;
  int mc_supported = 0;
  if (ok && (v.DeviceGetAttribute(&mc_supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) !=
                 CUDA_SUCCESS || !mc_supported))
    ok = 0;
  CUmulticastObjectProp mprop;
  memset(&mprop, 0, sizeof(mprop));
  mprop.numDevices = (unsigned)nranks;
  mprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  if (ok) {
    mprop.size = kNvlsHeaderBytes + a->data_bytes;
    if (v.MulticastGetGranularity(&gran, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED) !=
            CUDA_SUCCESS || gran == 0)
      ok = 0;
  }
  if (ok) {
    a->size = (kNvlsHeaderBytes + a->data_bytes + gran - 1) / gran * gran;
    mprop.size = a->size;
  }
  if (vote_sum(nccl_comm, ok) != nranks) {
    nvls_arena_destroy(a);
    set_last_error("NVLS: multicast memory is not available on every rank");
    return B200_UNAVAILABLE;
  }
  // ---- phase 2: rank 0 creates the multicast object and hands its fd to the others
  struct Slot {
    char name[96];
    int ok;
    int pad[7];
  };
  static_assert(sizeof(Slot) == 128, "slot size");
  Slot mine;
  memset(&mine, 0, sizeof(mine));
  int listen_sock = -1, mc_fd = -1;
  if (rank == 0) {
    int good = v.MulticastCreate(&a->mc, &mprop) == CUDA_SUCCESS;
    if (good) a->mc_valid = true;
    if (good && v.MemExportToShareableHandle(&mc_fd, a->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR,
                                             0) != CUDA_SUCCESS)
      good = 0;
    if (good) {
      static int counter = 0;
      snprintf(mine.name, sizeof(mine.name), "b200tf-nvls-%d-%d", (int)getpid(), counter++);
      listen_sock = socket(AF_UNIX, SOCK_STREAM, 0);
      struct sockaddr_un addr;
      const socklen_t len = abstract_addr(mine.name, &addr);
      if (listen_sock < 0 || bind(listen_sock, (struct sockaddr*)&addr, len) != 0 ||
          listen(listen_sock, nranks) != 0)
        good = 0;
    }
    mine.ok = good;
  } else {
    mine.ok = 1;
  }
  Slot* dev_slots = nullptr;
  Slot host_slots[8];
  memset(host_slots, 0, sizeof(host_slots));
  int exchanged = cudaMalloc(&dev_slots, sizeof(Slot) * 9) == cudaSuccess;
  if (exchanged) {
    cudaMemcpy(dev_slots + 8, &mine, sizeof(Slot), cudaMemcpyHostToDevice);
    exchanged = b200_nccl_all_gather_bytes(dev_slots + 8, dev_slots, sizeof(Slot), nccl_comm,
                                           nullptr) == B200_OK &&
                cudaDeviceSynchronize() == cudaSuccess &&
                cudaMemcpy(host_slots, dev_slots, sizeof(Slot) * nranks, cudaMemcpyDeviceToHost) ==
                    cudaSuccess;
  }
  if (dev_slots) cudaFree(dev_slots);
  cudaGetLastError();
  const bool root_ok = exchanged && host_slots[0].ok != 0;
  int step = root_ok ? 1 : 0;
  if (root_ok) {
    if (rank == 0) {
      for (int p = 1; p < nranks && step; ++p) {
        const int c = accept(listen_sock, nullptr, nullptr);
        if (c < 0 || send_fd(c, mc_fd, 1) != 0) step = 0;
        if (c >= 0) close(c);
      }
    } else {
      const int sock = socket(AF_UNIX, SOCK_STREAM, 0);
      struct sockaddr_un addr;
      const socklen_t len = abstract_addr(host_slots[0].name, &addr);
      int status = 0, fd = -1;
      if (sock < 0 || connect(sock, (struct sockaddr*)&addr, len) != 0 ||
          (fd = recv_fd(sock, &status)) < 0 || status != 1)
        step = 0;
      if (sock >= 0) close(sock);
      if (step) {
        if (v.MemImportFromShareableHandle(&a->mc, (void*)(uintptr_t)fd,
                                           CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS)
          a->mc_valid = true;
        else
          step = 0;
      }
      if (fd >= 0) close(fd);
    }
  }
  if (listen_sock >= 0) close(listen_sock);
  if (mc_fd >= 0) close(mc_fd);
  // ---- phase 3: add this device, then (after everybody has) bind local memory and map both views
  if (step && v.MulticastAddDevice(a->mc, dev) != CUDA_SUCCESS) step = 0;
  if (vote_sum(nccl_comm, step) != nranks) {
    nvls_arena_destroy(a);
    set_last_error("NVLS: could not share the multicast object with every rank");
    return B200_UNAVAILABLE;
  }
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = a->device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemAccessDesc access;
  memset(&access, 0, sizeof(access));
  access.location = prop.location;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  int mapped = 1;
  if (v.MemCreate(&a->mem, a->size, &prop, 0) != CUDA_SUCCESS) mapped = 0;
  if (mapped) a->mem_valid = true;
  if (mapped && v.MulticastBindMem(a->mc, 0, a->mem, 0, a->size, 0) != CUDA_SUCCESS) mapped = 0;
  if (mapped) a->bound = true;
  if (mapped && v.MemAddressReserve(&a->uc_va, a->size, gran, 0, 0) != CUDA_SUCCESS) mapped = 0;
  if (mapped && v.MemMap(a->uc_va, a->size, 0, a->mem, 0) != CUDA_SUCCESS) mapped = 0;
  if (mapped) a->uc_mapped = true;
  if (mapped && v.MemSetAccess(a->uc_va, a->size, &access, 1) != CUDA_SUCCESS) mapped = 0;
  if (mapped && v.MemAddressReserve(&a->mc_va, a->size, gran, 0, 0) != CUDA_SUCCESS) mapped = 0;
  if (mapped && v.MemMap(a->mc_va, a->size, 0, a->mc, 0) != CUDA_SUCCESS) mapped = 0;
  if (mapped) a->mc_mapped = true;
  if (mapped && v.MemSetAccess(a->mc_va, a->size, &access, 1) != CUDA_SUCCESS) mapped = 0;
  if (mapped && cudaMemset(reinterpret_cast<void*>(a->uc_va), 0, a->size) != cudaSuccess) mapped = 0;
  if (mapped && cudaDeviceSynchronize() != cudaSuccess) mapped = 0;
  cudaGetLastError();
  if (vote_sum(nccl_comm, mapped) != nranks) {
    nvls_arena_destroy(a);
    set_last_error("NVLS: binding / mapping the multicast memory failed on at least one rank");
    return B200_UNAVAILABLE;
  }
  *out = a;
  return B200_OK;
}

void* nvls_arena_data(NvlsArena* a) {
  return reinterpret_cast<char*>(a->uc_va) + kNvlsHeaderBytes;
}
size_t nvls_arena_bytes(NvlsArena* a) { return a->data_bytes; }

int nvls_all_reduce(NvlsArena* a, int dtype, size_t offset_bytes, long long count, int average,
                    int max_ctas, cudaStream_t stream) {
  if (count == 0) return B200_OK;
  const int per_vec = dtype == B200_DT_FLOAT ? 4 : 8;
  const long long nvec = (count + per_vec - 1) / per_vec;  // the arena is padded to 256 bytes
  const long long slice = (nvec + a->nranks - 1) / a->nranks;
  int ctas = max_ctas > 0 ? max_ctas : 128;
  if (ctas > kNvlsMaxCtas) ctas = kNvlsMaxCtas;
  long long useful = (slice + kNvlsThreads - 1) / kNvlsThreads;
  if (useful < 1) useful = 1;
  if (ctas > useful) ctas = (int)useful;
  char* mc = reinterpret_cast<char*>(a->mc_va);
  char* uc = reinterpret_cast<char*>(a->uc_va);
  const float scale = average ? 1.0f / (float)a->nranks : 1.0f;
  cudaError_t e;
  if (dtype == B200_DT_FLOAT)
    e = launch_pdl(nvls_all_reduce_kernel<false>, dim3(ctas), dim3(kNvlsThreads), 0, stream,
                   mc + kNvlsHeaderBytes + offset_bytes, reinterpret_cast<uint32_t*>(mc),
                   reinterpret_cast<const uint32_t*>(uc),
                   reinterpret_cast<uint32_t*>(uc + kNvlsCallsOffset), nvec, slice, a->rank,
                   a->nranks, scale);
  else
    e = launch_pdl(nvls_all_reduce_kernel<true>, dim3(ctas), dim3(kNvlsThreads), 0, stream,
                   mc + kNvlsHeaderBytes + offset_bytes, reinterpret_cast<uint32_t*>(mc),
                   reinterpret_cast<const uint32_t*>(uc),
                   reinterpret_cast<uint32_t*>(uc + kNvlsCallsOffset), nvec, slice, a->rank,
                   a->nranks, scale);
  if (e != cudaSuccess) {
    set_last_error("nvls_all_reduce launch: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return B200_INTERNAL;
  }
  note_launch();
  note_collective(true);
  return check_launch("nvls_all_reduce");
}

}  // namespace b200

extern "C" int b200_nvls_supported(void) {
  // 1: the current device and driver can put memory behind an NVSwitch multicast object (the
  // per-device half of nvls_arena_create's first vote); B200TF_NVLS=0 answers 0 as well.
  const char* nv = getenv("B200TF_NVLS");
  if (nv != nullptr && std::strcmp(nv, "0") == 0) return 0;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) {
    cudaGetLastError();
    return 0;
  }
  const b200::VmmApi& v = b200::vmm();
  if (!v.ok) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int mc = 0;
  if (v.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS)
    return 0;
  return mc ? 1 : 0;
}
