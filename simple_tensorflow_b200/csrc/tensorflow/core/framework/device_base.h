// DeviceBase / DeviceContext / GpuDeviceInfo -- subset of the reference's
// core/framework/device_base.h:59-224.
#ifndef B200TF_CORE_FRAMEWORK_DEVICE_BASE_H_
#define B200TF_CORE_FRAMEWORK_DEVICE_BASE_H_

#include <functional>
#include <string>

#include "tensorflow/core/framework/allocator.h"
#include "tensorflow/core/framework/tensor.h"
#include "tensorflow/stream_executor/stream.h"

namespace tensorflow {
namespace gpu = ::perftools::gputools;

class Device;
typedef std::function<void(const Status&)> StatusCallback;

// device_base.h:59-100: per-op device context; for GPU it names the stream the op enqueues on
// and implements the host<->device tensor copies used for feeds and fetches.
class DeviceContext {
 public:
  virtual ~DeviceContext() {}
  virtual gpu::Stream* stream() const { return nullptr; }
  virtual void CopyCPUTensorToDevice(const Tensor* cpu_tensor, Device* device,
                                     Tensor* device_tensor, StatusCallback done) const {
    done(errors::Internal("Unrecognized device type in CPU-to-device Copy"));
  }
  virtual void CopyDeviceTensorToCPU(const Tensor* device_tensor, const std::string& tensor_name,
                                     Device* device, Tensor* cpu_tensor, StatusCallback done) {
    done(errors::Internal("Unrecognized device type in device-to-CPU Copy"));
  }
};

class DeviceBase {
 public:
  virtual ~DeviceBase() {}
  // device_base.h:130-145
  struct GpuDeviceInfo {
    gpu::Stream* stream = nullptr;
    DeviceContext* default_context = nullptr;
    int gpu_id = -1;
  };
  void set_tensorflow_gpu_device_info(GpuDeviceInfo* g) { gpu_device_info_ = g; }
  const GpuDeviceInfo* tensorflow_gpu_device_info() const { return gpu_device_info_; }
  virtual Allocator* GetAllocator(AllocatorAttributes attr) = 0;

 private:
  GpuDeviceInfo* gpu_device_info_ = nullptr;
};

}  // namespace tensorflow
#endif
