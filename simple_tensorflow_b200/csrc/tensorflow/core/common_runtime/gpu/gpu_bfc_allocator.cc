#include "tensorflow/core/common_runtime/gpu/gpu_bfc_allocator.h"

#include <algorithm>

#include "b200_ops.h"

namespace tensorflow {

static size_t RoundUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

GPUBFCAllocator::GPUBFCAllocator(int device_id, size_t total_memory, const std::string& name)
    : device_id_(device_id), memory_limit_(total_memory), name_(name),
      next_region_bytes_(256u << 20) {}

GPUBFCAllocator::~GPUBFCAllocator() {
  b200_set_device(device_id_);
  for (auto& r : regions_) b200_free(r.first);
}

void GPUBFCAllocator::Unref() {
  if (refs_.fetch_sub(1, std::memory_order_acq_rel) == 1) delete this;
}

void GPUBFCAllocator::InsertFree(char* ptr, size_t size) { free_by_size_.insert({size, ptr}); }
void GPUBFCAllocator::RemoveFree(char* ptr, size_t size) { free_by_size_.erase({size, ptr}); }

bool GPUBFCAllocator::Extend(size_t rounded_bytes) {
  size_t want = std::max(next_region_bytes_, RoundUp(rounded_bytes, 2u << 20));
  if (memory_limit_ && stats_.bytes_reserved + want > memory_limit_) {
    want = RoundUp(rounded_bytes, 2u << 20);
    if (stats_.bytes_reserved + want > memory_limit_) return false;
  }
  void* p = nullptr;
  b200_set_device(device_id_);
  while (b200_malloc(&p, want) != 0 || p == nullptr) {
    // back off like BFCAllocator::Extend does (bfc_allocator.cc: "try 90% of the size")
    if (want <= RoundUp(rounded_bytes, 2u << 20)) return false;
    want = std::max(RoundUp(rounded_bytes, 2u << 20), RoundUp(want / 2, 2u << 20));
  }
  regions_.push_back({static_cast<char*>(p), want});
  stats_.bytes_reserved += want;
  next_region_bytes_ = std::min<size_t>(next_region_bytes_ * 2, size_t(16) << 30);
  chunks_[static_cast<char*>(p)] = Chunk{static_cast<char*>(p), want, false};
  InsertFree(static_cast<char*>(p), want);
  return true;
}

void* GPUBFCAllocator::AllocateRaw(size_t alignment, size_t num_bytes) {
  void* p = AllocateLocked(alignment, num_bytes);
  if (p != nullptr) refs_.fetch_add(1, std::memory_order_relaxed);  // released by DeallocateRaw
  return p;
}

void* GPUBFCAllocator::AllocateLocked(size_t /*alignment*/, size_t num_bytes) {
  if (num_bytes == 0) return nullptr;
  const size_t rounded = RoundUp(num_bytes, kMinAllocationSize);
  std::lock_guard<std::mutex> l(mu_);
  auto it = free_by_size_.lower_bound({rounded, nullptr});
  if (it == free_by_size_.end()) {
    if (!Extend(rounded)) return nullptr;
    it = free_by_size_.lower_bound({rounded, nullptr});
    if (it == free_by_size_.end()) return nullptr;
  }
  char* ptr = it->second;
  const size_t size = it->first;
  free_by_size_.erase(it);
  Chunk& c = chunks_[ptr];
  c.in_use = true;
  // split when the remainder is worth keeping (bfc_allocator.cc SplitChunk)
  if (size - rounded >= kMinAllocationSize) {
    c.size = rounded;
    char* rest = ptr + rounded;
    chunks_[rest] = Chunk{rest, size - rounded, false};
    InsertFree(rest, size - rounded);
  }
  stats_.num_allocs++;
  stats_.bytes_in_use += c.size;
  stats_.max_bytes_in_use = std::max(stats_.max_bytes_in_use, stats_.bytes_in_use);
  return ptr;
}

void GPUBFCAllocator::BeginPin() {
  std::lock_guard<std::mutex> l(mu_);
  pin_mode_ = true;
  pinned_.clear();
}
void GPUBFCAllocator::EndPin(std::vector<void*>* pinned) {
  std::lock_guard<std::mutex> l(mu_);
  pin_mode_ = false;
  pinned->swap(pinned_);
  pinned_.clear();
}

void GPUBFCAllocator::DeallocateRaw(void* p) {
  if (p == nullptr) return;
  {
    std::lock_guard<std::mutex> l(mu_);
    if (pin_mode_) {
      pinned_.push_back(p);  // stays allocated until the captured graph is destroyed
      return;
    }
  }
  if (DeallocateLocked(p)) Unref();  // may delete this: the lock is already released
}

bool GPUBFCAllocator::DeallocateLocked(void* p) {
  std::lock_guard<std::mutex> l(mu_);
  auto it = chunks_.find(static_cast<char*>(p));
  if (it == chunks_.end() || !it->second.in_use) return false;
  it->second.in_use = false;
  stats_.bytes_in_use -= it->second.size;
  // coalesce with the next chunk, if free and in the same region (contiguous address)
  auto next = std::next(it);
  if (next != chunks_.end() && !next->second.in_use &&
      it->second.ptr + it->second.size == next->second.ptr) {
    bool region_start = false;
    for (auto& r : regions_) region_start |= (r.first == next->second.ptr);
    if (!region_start) {
      RemoveFree(next->second.ptr, next->second.size);
      it->second.size += next->second.size;
      chunks_.erase(next);
    }
  }
  // coalesce with the previous chunk
  if (it != chunks_.begin()) {
    auto prev = std::prev(it);
    bool region_start = false;
    for (auto& r : regions_) region_start |= (r.first == it->second.ptr);
    if (!region_start && !prev->second.in_use &&
        prev->second.ptr + prev->second.size == it->second.ptr) {
      RemoveFree(prev->second.ptr, prev->second.size);
      prev->second.size += it->second.size;
      chunks_.erase(it);
      it = prev;
    }
  }
  InsertFree(it->second.ptr, it->second.size);
  return true;
}

void GPUBFCAllocator::GetStats(AllocatorStats* stats) {
  std::lock_guard<std::mutex> l(mu_);
  *stats = stats_;
}

GPUHostAllocator* GPUHostAllocator::Process() {
  static GPUHostAllocator* a = new GPUHostAllocator;  // leaked on purpose: process lifetime
  return a;
}
GPUHostAllocator::~GPUHostAllocator() {
  for (auto& kv : free_) b200_host_free(kv.second);
  for (auto& kv : live_) b200_host_free(kv.first);  // nothing may still use them at this point
}
void* GPUHostAllocator::AllocateRaw(size_t, size_t num_bytes) {
  if (num_bytes == 0) return nullptr;
  size_t bucket = 4096;
  while (bucket < num_bytes) bucket <<= 1;
  std::lock_guard<std::mutex> l(mu_);
  auto it = free_.find(bucket);
  void* p = nullptr;
  if (it != free_.end()) {
    p = it->second;
    free_.erase(it);
  } else if (b200_host_malloc(&p, bucket) != 0) {
    return nullptr;
  }
  live_[p] = bucket;
  return p;
}
void GPUHostAllocator::DeallocateRaw(void* ptr) {
  if (!ptr) return;
  std::lock_guard<std::mutex> l(mu_);
  auto it = live_.find(ptr);
  if (it == live_.end()) return;
  free_.insert({it->second, ptr});
  live_.erase(it);
}

}  // namespace tensorflow
