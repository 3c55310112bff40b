// Allocator / AllocatorAttributes -- subset of the reference's core/framework/allocator.h:67-302.
#ifndef B200TF_CORE_FRAMEWORK_ALLOCATOR_H_
#define B200TF_CORE_FRAMEWORK_ALLOCATOR_H_

#include <cstdlib>
#include <string>

#include "tensorflow/core/framework/types.h"

namespace tensorflow {

struct AllocatorStats {
  int64 num_allocs = 0;
  int64 bytes_in_use = 0;
  int64 max_bytes_in_use = 0;
  int64 bytes_reserved = 0;
};

class Allocator {
 public:
  // allocator.h:70-76 uses 32 (64 with AVX-512); device arenas round up to 256 themselves.
  static constexpr size_t kAllocatorAlignment = 64;
  virtual ~Allocator() {}
  virtual std::string Name() = 0;
  virtual void* AllocateRaw(size_t alignment, size_t num_bytes) = 0;
  virtual void DeallocateRaw(void* ptr) = 0;
  virtual void GetStats(AllocatorStats* stats) { *stats = AllocatorStats(); }
};

// allocator.h:239-283: one bit per property, merged with |=.
struct AllocatorAttributes {
  void set_on_host(bool v) { value = v ? (value | 0x1) : (value & ~0x1u); }
  bool on_host() const { return value & 0x1; }
  void set_gpu_compatible(bool v) { value = v ? (value | 0x4) : (value & ~0x4u); }
  bool gpu_compatible() const { return value & 0x4; }
  void Merge(AllocatorAttributes other) { value |= other.value; }
  uint32 value = 0;
};

// Plain host allocator (posix_memalign), the role of cpu_allocator() (allocator.cc).
class CPUAllocator : public Allocator {
 public:
  std::string Name() override { return "cpu"; }
  void* AllocateRaw(size_t alignment, size_t num_bytes) override {
    void* p = nullptr;
    if (num_bytes == 0) return nullptr;
    if (alignment < sizeof(void*)) alignment = sizeof(void*);
    if (posix_memalign(&p, alignment, num_bytes) != 0) return nullptr;
    return p;
  }
  void DeallocateRaw(void* ptr) override { free(ptr); }
};
inline Allocator* cpu_allocator() {
  static CPUAllocator* a = new CPUAllocator;
  return a;
}

}  // namespace tensorflow
#endif
