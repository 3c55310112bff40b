// perftools::gputools::Stream / Event / DeviceMemoryBase -- the slice of the reference's
// StreamExecutor surface that its GPU device and kernels use (stream_executor/stream.h:116,189,
// 214,1482-1531,1591; device_memory.h:47), implemented as a thin veneer over the C ABI of
// libb200tf.so (include/b200_ops.h).  No CUDA headers are needed above this line.
#ifndef B200TF_STREAM_EXECUTOR_STREAM_H_
#define B200TF_STREAM_EXECUTOR_STREAM_H_

#include <cstddef>
#include <cstdint>
#include <functional>

#include "b200_ops.h"

namespace perftools {
namespace gputools {

class DeviceMemoryBase {
 public:
  explicit DeviceMemoryBase(void* opaque = nullptr, uint64_t size = 0)
      : opaque_(opaque), size_(size) {}
  void* opaque() { return opaque_; }
  const void* opaque() const { return opaque_; }
  uint64_t size() const { return size_; }
  bool is_null() const { return opaque_ == nullptr; }

 private:
  void* opaque_;
  uint64_t size_;
};

class Event {
 public:
  Event() : handle_(nullptr) {}
  ~Event() {
    if (handle_) b200_event_destroy(handle_);
  }
  bool Init() { return b200_event_create(&handle_) == 0; }
  // Event::PollForStatus: kComplete / kPending / kError
  enum class Status { kComplete, kPending, kError };
  Status PollForStatus() {
    const int r = b200_event_query(handle_);
    return r == 0 ? Status::kComplete : (r == 1 ? Status::kPending : Status::kError);
  }
  void* handle() const { return handle_; }

 private:
  void* handle_;
};

// Error convention of the reference: Then* return *this and latch failure in ok()
// (kernels then map !ok() to errors::Internal, e.g. core/kernels/matmul_op.cc:196-201).
class Stream {
 public:
  Stream() : handle_(nullptr), ok_(false), owned_(false) {}
  // Wrap an existing CUstream (e.g. the framework embedding us already owns one).
  explicit Stream(void* existing) : handle_(existing), ok_(true), owned_(false) {}
  ~Stream() {
    if (owned_ && handle_) b200_stream_destroy(handle_);
  }
  Stream& Init(bool high_priority = false) {
    ok_ = (high_priority ? b200_stream_create_with_priority(&handle_, 1)
                         : b200_stream_create(&handle_)) == 0;
    owned_ = ok_;
    return *this;
  }
  bool ok() const { return ok_; }
  void* cuda_stream() const { return handle_; }  // what kernels pass to the b200_* ops

  Stream& ThenMemcpyH2D(DeviceMemoryBase* gpu_dst, const void* host_src, uint64_t size) {
    return Latch(b200_memcpy_h2d_async(gpu_dst->opaque(), host_src, size, handle_));
  }
  Stream& ThenMemcpyD2H(void* host_dst, const DeviceMemoryBase& gpu_src, uint64_t size) {
    return Latch(b200_memcpy_d2h_async(host_dst, gpu_src.opaque(), size, handle_));
  }
  Stream& ThenMemcpyD2D(DeviceMemoryBase* gpu_dst, const DeviceMemoryBase& gpu_src,
                        uint64_t size) {
    return Latch(b200_memcpy_d2d_async(gpu_dst->opaque(), gpu_src.opaque(), size, handle_));
  }
  // The reference's overload set (stream_executor/stream.h:1482-1529): direction from the types.
  Stream& ThenMemcpy(void* host_dst, const DeviceMemoryBase& gpu_src, uint64_t size) {
    return ThenMemcpyD2H(host_dst, gpu_src, size);
  }
  Stream& ThenMemcpy(DeviceMemoryBase* gpu_dst, const void* host_src, uint64_t size) {
    return ThenMemcpyH2D(gpu_dst, host_src, size);
  }
  Stream& ThenMemcpy(DeviceMemoryBase* gpu_dst, const DeviceMemoryBase& gpu_src, uint64_t size) {
    return ThenMemcpyD2D(gpu_dst, gpu_src, size);
  }
  // stream_executor/stream.h:1624: run `callback` on a driver thread once the work enqueued so
  // far has completed (the callback must not call into the stream).
  Stream& ThenDoHostCallback(std::function<void()> callback) {
    auto* heap = new std::function<void()>(std::move(callback));
    const int rc = b200_stream_add_host_callback(handle_, &Stream::RunHostCallback, heap);
    if (rc != 0) delete heap;
    return Latch(rc);
  }
  // (ThenLaunch has no counterpart: kernels are enqueued through the b200_* entry points, which
  // take this stream's handle -- cuda_stream() -- as their last argument.)
  Stream& ThenMemZero(DeviceMemoryBase* location, uint64_t size) {
    return Latch(b200_memset_async(location->opaque(), 0, size, handle_));
  }
  Stream& ThenRecordEvent(Event* event) {
    return Latch(b200_event_record(event->handle(), handle_));
  }
  Stream& ThenWaitFor(Event* event) {
    return Latch(b200_stream_wait_event(handle_, event->handle()));
  }
  bool BlockHostUntilDone() {
    Latch(b200_stream_synchronize(handle_));
    return ok_;
  }

 private:
  static void RunHostCallback(void* arg) {
    auto* fn = static_cast<std::function<void()>*>(arg);
    (*fn)();
    delete fn;
  }
  Stream& Latch(int rc) {
    if (rc != 0) ok_ = false;
    return *this;
  }
  void* handle_;
  bool ok_;
  bool owned_;
};

}  // namespace gputools
}  // namespace perftools
#endif
