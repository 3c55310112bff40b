// GPUBFCAllocator -- a best-fit-with-coalescing arena over b200_malloc'd regions, playing the
// role of the reference's core/common_runtime/gpu/gpu_bfc_allocator.{h,cc} +
// core/common_runtime/bfc_allocator.{h,cc}.
//
// Same contract: AllocateRaw never calls cudaMalloc on the steady-state path, chunks are split
// and coalesced, and a freed chunk may be handed out again IMMEDIATELY -- safe because every
// kernel of a device is enqueued on its single compute stream, so reuse is stream-ordered
// (the argument of common_runtime/gpu/gpu_device.cc:266-271).  Regions grow by doubling
// (bfc_allocator.cc Extend), each region is one b200_malloc.
#ifndef B200TF_CORE_COMMON_RUNTIME_GPU_GPU_BFC_ALLOCATOR_H_
#define B200TF_CORE_COMMON_RUNTIME_GPU_GPU_BFC_ALLOCATOR_H_

#include <atomic>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "tensorflow/core/framework/allocator.h"

namespace tensorflow {

class GPUBFCAllocator : public Allocator {
 public:
  // total_memory: cap on the sum of regions (0 = no cap beyond the device).
  GPUBFCAllocator(int device_id, size_t total_memory, const std::string& name);
  ~GPUBFCAllocator() override;
  // Lifetime: the device that created the arena holds one reference and every live allocation
  // holds one, so a Tensor that outlives its session (TF_DeleteSession before TF_DeleteTensor is
  // legal in the reference's C API) still deallocates into a live allocator; the regions are
  // returned to the driver when the last of them goes.  The device calls Unref() instead of delete.
  void Unref();
  // Pin mode (step-level CUDA graphs): while a step is being captured, nothing it frees may be
  // handed out again -- the captured kernels keep the addresses -- so DeallocateRaw parks the
  // pointers; EndPin returns them and the graph's owner frees them when the graph dies.
  void BeginPin();
  void EndPin(std::vector<void*>* pinned);
  std::string Name() override { return name_; }
  void* AllocateRaw(size_t alignment, size_t num_bytes) override;
  void DeallocateRaw(void* ptr) override;
  void GetStats(AllocatorStats* stats) override;

  static constexpr size_t kMinAllocationSize = 256;  // bfc_allocator.h kMinAllocationSize

 private:
  struct Chunk {
    char* ptr;
    size_t size;
    bool in_use;
  };
  bool Extend(size_t rounded_bytes);
  void* AllocateLocked(size_t alignment, size_t num_bytes);
  bool DeallocateLocked(void* ptr);
  void InsertFree(char* ptr, size_t size);
  void RemoveFree(char* ptr, size_t size);

  const int device_id_;
  const size_t memory_limit_;
  const std::string name_;
  std::mutex mu_;
  std::map<char*, Chunk> chunks_;                  // all chunks by address
  std::set<std::pair<size_t, char*>> free_by_size_;  // free chunks, best-fit lookup
  std::vector<std::pair<char*, size_t>> regions_;
  size_t next_region_bytes_;
  AllocatorStats stats_;
  std::atomic<long long> refs_{1};
  bool pin_mode_ = false;
  std::vector<void*> pinned_;
};

// Pinned host memory for feeds/fetches (the role of PoolAllocator + CUDAHostAllocator,
// common_runtime/gpu/pool_allocator.h): size-bucketed free lists over b200_host_malloc.
class GPUHostAllocator : public Allocator {
 public:
  // The process-lifetime instance every device and the C API share (the reference's
  // ProcessState::GetCUDAHostAllocator, common_runtime/gpu/process_state.cc): fetched tensors
  // may outlive the session that produced them.  Never destroyed.
  static GPUHostAllocator* Process();
  ~GPUHostAllocator() override;
  std::string Name() override { return "cuda_host_bfc"; }
  void* AllocateRaw(size_t alignment, size_t num_bytes) override;
  void DeallocateRaw(void* ptr) override;

 private:
  std::mutex mu_;
  std::multimap<size_t, void*> free_;
  std::map<void*, size_t> live_;
};

}  // namespace tensorflow
#endif
