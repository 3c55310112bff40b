// All-reduce over NVLink peer memory (one process per GPU, buffers mapped with CUDA IPC).
//
// Why it exists (profiles/r01_notes.md): NCCL's all-reduce of the 12.6 MB gradient arena costs
// 58 / 77 us on 4 / 8 B200 and is latency-bound (27 us for 4 KB); run beside the backward pass
// its 24-32 CTAs scatter over as many TPCs and break up the CTA pairs of the persistent tcgen05
// GEMMs.  This kernel is a two-shot all-reduce (reduce-scatter by the slice owner, then
// all-gather) written for exactly that situation: LDG.128 peer loads (775 GB/s per GPU, 1 us
// latency, L2-bypass), three flag barriers in peer memory, and CTAs small enough (256 threads,
// <= 64 registers, no shared memory) to be resident on an SM next to a GEMM CTA, so the exchange
// can run under the backward pass without taking SMs from it.  Sums run in rank order in fp32
// and only the slice owner reduces a slice, so every rank ends with identical bits.
//
// Reference role: the gradient exchange the reference's tower pattern does with _Send/_Recv +
// AddN (core/kernels/aggregate_ops.cc:153-176) and D2D copies (common_runtime/gpu/gpu_util.cc:190-250).
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "b200_internal.h"

namespace b200 {
// nvls_allreduce.cu
struct NvlsArena;
int nvls_arena_create(void* nccl_comm, int rank, int nranks, size_t data_bytes, NvlsArena** out);
void nvls_arena_destroy(NvlsArena* a);
void* nvls_arena_data(NvlsArena* a);
size_t nvls_arena_bytes(NvlsArena* a);
int nvls_all_reduce(NvlsArena* a, int dtype, size_t offset_bytes, long long count, int average,
                    int max_ctas, cudaStream_t stream);
namespace {

constexpr int kMaxRanks = 8;
constexpr int kMaxCtas = 512;
constexpr size_t kFlagBytes = (size_t)kMaxRanks * kMaxCtas * sizeof(uint32_t);  // 16 KB
constexpr size_t kHeaderBytes = 64 << 10;  // flags live in the first 64 KB of every rank's buffer
constexpr size_t kEpochOffset = 32 << 10;  // per-CTA call counters of the owning rank (local only)
// 256 threads x <= 64 registers: a CTA of this kernel fits on an SM BESIDE a resident tcgen05 GEMM
// CTA (192 threads x <= 192 registers, ~220 KB smem), so on the collective stream it needs no
// free SMs and takes none from the GEMMs -- it only shares their issue slots.
constexpr int kThreads = 256;

struct PeerTable {
  char* base[kMaxRanks];  // every rank's buffer as mapped into THIS process (base[rank] = own)
  int rank, nranks;
};

const char* g_backend = "none";  // what the last arena created in this process runs on

struct PeerArena {
  NvlsArena* nvls = nullptr;  // non-null: the arena lives in NVSwitch multicast memory instead
  PeerTable table;
  size_t data_bytes = 0;
  uint32_t epoch = 0;  // barrier values used so far (all ranks issue the same call sequence)
  int device = 0;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data: never keep it in L1 (the same addresses are rewritten every step by another GPU)
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// CTA b of every rank meets CTA b of every other rank.  Flags only grow (value = running epoch),
// so nothing is ever reset.  The leading __syncthreads + release make this CTA's earlier global
// writes visible to whoever observes the flag; the acquire + trailing __syncthreads order the
// peers' writes before this CTA's later reads.
__device__ __forceinline__ void peer_barrier(const PeerTable& t, uint32_t value) {
  __syncthreads();
  if ((int)threadIdx.x < t.nranks) {
    const int peer = threadIdx.x;
    uint32_t* theirs = reinterpret_cast<uint32_t*>(t.base[peer]) + t.rank * kMaxCtas + blockIdx.x;
    st_release_sys(theirs, value);
    const uint32_t* mine =
        reinterpret_cast<const uint32_t*>(t.base[t.rank]) + peer * kMaxCtas + blockIdx.x;
    while (ld_acquire_sys(mine) < value) {
    }
  }
  __syncthreads();
}

// In place over `count` floats at byte offset `offset` of every rank's data region.
// Slice q = vectors [q*L, (q+1)*L) belongs to rank q.  CTA b touches the same vector indices of a
// slice on every rank, which is what makes the CTA-to-CTA barriers sufficient.
template <int NR>
__global__ void __launch_bounds__(kThreads, 4)
peer_all_reduce_kernel(const __grid_constant__ PeerTable t, size_t offset, long long nvec, long long slice_vecs,
                       float scale) {
  pdl_prologue();
  // The barrier epoch lives in device memory (one word per CTA in this rank's header, never
  // touched by peers): a kernel argument would be frozen into a captured CUDA graph, and every
  // replay must use fresh, growing flag values.  All ranks issue the same call sequence with the
  // same grid, so CTA b's counter agrees across ranks.
  __shared__ uint32_t epoch_sm;
  if (threadIdx.x == 0) {
    uint32_t* calls = reinterpret_cast<uint32_t*>(t.base[t.rank] + kEpochOffset) + blockIdx.x;
    epoch_sm = *calls;
    *calls = epoch_sm + 3;
  }
  __syncthreads();
  const uint32_t epoch = epoch_sm;
  const int rank = t.rank;
  // (pointers are re-derived from the parameter table: indexing a local array by rank would
  // put it on the stack)
  auto buf = [&](int p) { return reinterpret_cast<float4*>(t.base[p] + kHeaderBytes + offset); };
  float4* const mine = buf(rank);
  // (0) every rank's producers have finished (each rank's kernel is stream-ordered behind them)
  peer_barrier(t, epoch + 1);
  // (1) reduce-scatter: the owner sums its slice over all ranks, in rank order.  U vectors per
  // thread are in flight at once: a peer load takes ~1 us, so bandwidth = bytes in flight / 1 us
  // (CTAs x 256 threads x U x (NR-1) x 16 B must reach ~1 MB to fill NVLink: 32-64 CTAs).
  constexpr int U = NR <= 2 ? 4 : (NR <= 4 ? 2 : 1);  // U x NR <= 8 vectors = 32 registers
  const long long lo = (long long)rank * slice_vecs;
  long long hi = lo + slice_vecs;
  if (hi > nvec) hi = nvec;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long v0 = lo + (long long)blockIdx.x * kThreads + threadIdx.x; v0 < hi;
       v0 += stride * U) {
    float4 x[U][NR];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * stride;
      if (v < hi) {
#pragma unroll
        for (int p = 0; p < NR; ++p) x[u][p] = p == rank ? mine[v] : ld_peer(buf(p) + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = v0 + u * stride;
      if (v < hi) {
        float4 s = x[u][0];
#pragma unroll
        for (int p = 1; p < NR; ++p) {
          s.x += x[u][p].x;
          s.y += x[u][p].y;
          s.z += x[u][p].z;
          s.w += x[u][p].w;
        }
        s.x *= scale;
        s.y *= scale;
        s.z *= scale;
        s.w *= scale;
        mine[v] = s;
      }
    }
  }
  // (2) every owner has published its slice
  peer_barrier(t, epoch + 2);
  // (3) all-gather: pull the other owners' slices (same per-CTA index pattern as in (1)); the
  // loads of all NR-1 peers (and G vectors of each) are issued before the first store
  constexpr int G = NR <= 2 ? 4 : (NR <= 4 ? 2 : 1);
  for (long long w0 = (long long)blockIdx.x * kThreads + threadIdx.x; w0 < slice_vecs;
       w0 += stride * G) {
    float4 y[G][NR];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long w = w0 + g * stride;
#pragma unroll
      for (int q = 1; q < NR; ++q) {
        const int p = (rank + q) % NR;  // start at a different peer on every rank
        const long long v = (long long)p * slice_vecs + w;
        if (w < slice_vecs && v < nvec) y[g][q] = ld_peer(buf(p) + v);
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const long long w = w0 + g * stride;
#pragma unroll
      for (int q = 1; q < NR; ++q) {
        const int p = (rank + q) % NR;
        const long long v = (long long)p * slice_vecs + w;
        if (w < slice_vecs && v < nvec) mine[v] = y[g][q];
      }
    }
  }
  // (4) nobody still reads my slice when this grid retires (the next step overwrites the arena)
  peer_barrier(t, epoch + 3);
}

}  // namespace
}  // namespace b200

using namespace b200;

#define PEER_CUDA(expr, what)                                         \
  do {                                                                \
    cudaError_t _e = (expr);                                          \
    if (_e != cudaSuccess) {                                          \
      set_last_error("%s: %s", what, cudaGetErrorString(_e));         \
      cudaGetLastError();                                             \
      return B200_INTERNAL;                                           \
    }                                                                 \
  } while (0)

int b200_peer_arena_create(void* nccl_comm, int rank, int nranks, size_t data_bytes, void** out) {
  *out = nullptr;
  if (nranks < 2 || nranks > kMaxRanks || rank < 0 || rank >= nranks) {
    set_last_error("b200_peer_arena_create: %d ranks unsupported (2..%d)", nranks, kMaxRanks);
    return B200_INVALID_ARGUMENT;
  }
  int rc = require_device("b200_peer_arena_create");
  if (rc) return rc;
  {
    // First choice: NVSwitch multicast memory (in-switch reduction).  The attempt is collective and
    // ends in a vote, so either every rank gets an NVLS arena or every rank falls through to the
    // peer-mapped arena below.  B200TF_NVLS=0 skips it.
    const char* nv = getenv("B200TF_NVLS");
    if (nv == nullptr || strcmp(nv, "0") != 0) {
      NvlsArena* n = nullptr;
      if (nvls_arena_create(nccl_comm, rank, nranks, data_bytes, &n) == B200_OK) {
        PeerArena* a = new PeerArena();
        memset(&a->table, 0, sizeof(a->table));
        cudaGetDevice(&a->device);
        a->nvls = n;
        a->data_bytes = nvls_arena_bytes(n);
        a->table.rank = rank;
        a->table.nranks = nranks;
        *out = a;
        g_backend = "nvls";
        return B200_OK;
      }
    }
  }
  // owns the arena record, this rank's buffer and the handle scratch until creation succeeded
  struct Guard {
    PeerArena* a = nullptr;
    char* local = nullptr;
    void* scratch = nullptr;
    void* vote = nullptr;
    ~Guard() {
      if (scratch) cudaFree(scratch);
      if (vote) cudaFree(vote);
      if (a) {
        for (int p = 0; p < a->table.nranks; ++p)
          if (p != a->table.rank && a->table.base[p]) cudaIpcCloseMemHandle(a->table.base[p]);
        delete a;
      }
      if (local) cudaFree(local);
      cudaGetLastError();
    }
  } guard;
  PeerArena* a = new PeerArena();
  guard.a = a;
  memset(&a->table, 0, sizeof(a->table));
  PEER_CUDA(cudaGetDevice(&a->device), "cudaGetDevice");
  a->data_bytes = (data_bytes + 255) / 256 * 256;
  a->table.rank = rank;
  a->table.nranks = nranks;
  const size_t total = kHeaderBytes + a->data_bytes;
  char* local = nullptr;
  cudaIpcMemHandle_t mine;
  int ok = cudaMalloc(&local, total) == cudaSuccess && (guard.local = local) != nullptr &&
           cudaMemset(local, 0, total) == cudaSuccess &&
           cudaIpcGetMemHandle(&mine, local) == cudaSuccess &&
           cudaDeviceSynchronize() == cudaSuccess;
  cudaGetLastError();
  // exchange the handles (and whether everybody got this far) through the communicator
  struct Slot {
    cudaIpcMemHandle_t handle;
    int ok;
    int pad[15];
  };
  static_assert(sizeof(Slot) == 128, "slot size");
  Slot host_slots[kMaxRanks];
  Slot my_slot;
  memset(&my_slot, 0, sizeof(my_slot));
  if (ok) my_slot.handle = mine;
  my_slot.ok = ok;
  Slot* dev_slots = nullptr;
  PEER_CUDA(cudaMalloc(&dev_slots, sizeof(Slot) * (kMaxRanks + 1)), "cudaMalloc(handles)");
  guard.scratch = dev_slots;
  PEER_CUDA(cudaMemcpy(dev_slots + kMaxRanks, &my_slot, sizeof(Slot), cudaMemcpyHostToDevice),
            "cudaMemcpy(handle)");
  rc = b200_nccl_all_gather_bytes(dev_slots + kMaxRanks, dev_slots, sizeof(Slot), nccl_comm, nullptr);
  if (rc == B200_OK) {
    PEER_CUDA(cudaDeviceSynchronize(), "handle exchange");
    PEER_CUDA(cudaMemcpy(host_slots, dev_slots, sizeof(Slot) * nranks, cudaMemcpyDeviceToHost),
              "cudaMemcpy(handles)");
  }
  bool all_ok = rc == B200_OK;
  for (int p = 0; all_ok && p < nranks; ++p) all_ok = host_slots[p].ok != 0;
  // map the peers; a rank that cannot must tell the others, hence a second vote
  int mapped = all_ok;
  for (int p = 0; mapped && p < nranks; ++p) {
    if (p == rank) {
      a->table.base[p] = local;
      continue;
    }
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, host_slots[p].handle, cudaIpcMemLazyEnablePeerAccess) !=
        cudaSuccess) {
      set_last_error("cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(cudaGetLastError()));
      mapped = 0;
      break;
    }
    a->table.base[p] = static_cast<char*>(ptr);
  }
  if (rc == B200_OK) {  // everybody reaches this collective whatever happened above
    float* vote = nullptr;
    PEER_CUDA(cudaMalloc(&vote, sizeof(float)), "cudaMalloc(vote)");
    guard.vote = vote;
    const float v = mapped ? 1.f : 0.f;
    PEER_CUDA(cudaMemcpy(vote, &v, sizeof(float), cudaMemcpyHostToDevice), "vote");
    rc = b200_nccl_all_reduce_sum(B200_DT_FLOAT, vote, vote, 1, nccl_comm, nullptr);
    float sum = 0.f;
    if (rc == B200_OK) {
      PEER_CUDA(cudaDeviceSynchronize(), "vote");
      PEER_CUDA(cudaMemcpy(&sum, vote, sizeof(float), cudaMemcpyDeviceToHost), "vote");
    }
    if (rc != B200_OK || (int)(sum + 0.5f) != nranks) mapped = 0;
  } else {
    mapped = 0;
  }
  if (!mapped) {
    set_last_error("b200_peer_arena_create: peer mapping unavailable on at least one rank");
    return B200_UNAVAILABLE;  // the guard releases everything
  }
  g_backend = "peer-ipc";
  guard.a = nullptr;      // success: the caller owns the arena (and through it the buffer)
  guard.local = nullptr;
  *out = a;
  return B200_OK;
}

int b200_peer_arena_destroy(void* arena) {
  PeerArena* a = static_cast<PeerArena*>(arena);
  if (!a) return B200_OK;
  if (a->nvls) {
    nvls_arena_destroy(a->nvls);
    delete a;
    return B200_OK;
  }
  cudaSetDevice(a->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < a->table.nranks; ++p)
    if (p != a->table.rank && a->table.base[p]) cudaIpcCloseMemHandle(a->table.base[p]);
  cudaFree(a->table.base[a->table.rank]);
  cudaGetLastError();
  delete a;
  return B200_OK;
}

const char* b200_peer_arena_backend(void) { return g_backend; }

void* b200_peer_arena_data(void* arena) {
  PeerArena* a = static_cast<PeerArena*>(arena);
  if (a && a->nvls) return nvls_arena_data(a->nvls);
  return a ? a->table.base[a->table.rank] + kHeaderBytes : nullptr;
}

size_t b200_peer_arena_bytes(void* arena) {
  PeerArena* a = static_cast<PeerArena*>(arena);
  return a ? a->data_bytes : 0;
}

int b200_peer_all_reduce(void* arena, int dtype, size_t offset_bytes, int64_t count, int average,
                         int max_ctas, void* stream) {
  PeerArena* a = static_cast<PeerArena*>(arena);
  // fp32 on both backends; bfloat16 only through the switch (multimem.ld_reduce accumulates in fp32)
  const bool dtype_ok = dtype == B200_DT_FLOAT || (a && a->nvls && dtype == B200_DT_BFLOAT16);
  const size_t es = dtype == B200_DT_FLOAT ? 4 : 2;
  if (!a || !dtype_ok || count < 0 || offset_bytes % 16 != 0 ||
      offset_bytes + (size_t)count * es > a->data_bytes) {
    set_last_error("b200_peer_all_reduce: bad arguments (fp32, or bf16 on the NVLS backend; "
                   "16-byte aligned offset, range inside the arena)");
    return B200_INVALID_ARGUMENT;
  }
  if (count == 0) return B200_OK;
  if (a->nvls)
    return nvls_all_reduce(a->nvls, dtype, offset_bytes, count, average, max_ctas, as_stream(stream));
  const int nr = a->table.nranks;
  const long long nvec = (count + 3) / 4;  // the arena is padded to 256 bytes: whole vectors
  const long long slice = (nvec + nr - 1) / nr;
  // Few CTAs keep the three barriers cheap (each CTA exchanges flags with every peer); 64 x 256
  // threads x 8 vectors in flight already cover the NVLink bandwidth-delay product.
  int ctas = max_ctas > 0 ? max_ctas : 64;
  if (ctas > kMaxCtas) ctas = kMaxCtas;
  long long useful = (slice + kThreads - 1) / kThreads;
  if (useful < 1) useful = 1;
  if (ctas > useful) ctas = (int)useful;
  const float scale = average ? 1.0f / (float)nr : 1.0f;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kThreads);
  cfg.stream = as_stream(stream);
  cudaError_t e = cudaSuccess;
#define PEER_LAUNCH(NR)                                                                      \
  case NR:                                                                                   \
    e = cudaLaunchKernelEx(&cfg, peer_all_reduce_kernel<NR>, a->table, offset_bytes, nvec,   \
                           slice, scale);                                                    \
    break;
  switch (nr) {
    PEER_LAUNCH(2)
    PEER_LAUNCH(3)
    PEER_LAUNCH(4)
    PEER_LAUNCH(5)
    PEER_LAUNCH(6)
    PEER_LAUNCH(7)
    PEER_LAUNCH(8)
    default:
      set_last_error("b200_peer_all_reduce: %d ranks", nr);
      return B200_INVALID_ARGUMENT;
  }
#undef PEER_LAUNCH
  if (e != cudaSuccess) {
    set_last_error("b200_peer_all_reduce launch: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return B200_INTERNAL;
  }
  note_launch();
  note_collective(true);
  return check_launch("b200_peer_all_reduce");
}
