// Device -- core/common_runtime/device.h:81-126 (Compute, Sync, name, device_type).
#ifndef B200TF_CORE_COMMON_RUNTIME_DEVICE_H_
#define B200TF_CORE_COMMON_RUNTIME_DEVICE_H_

#include <string>

#include "tensorflow/core/framework/device_base.h"
#include "tensorflow/core/framework/op_kernel.h"

namespace tensorflow {

class Device : public DeviceBase {
 public:
  Device(const std::string& name, const std::string& device_type)
      : name_(name), device_type_(device_type) {}
  ~Device() override {}
  const std::string& name() const { return name_; }
  const std::string& device_type() const { return device_type_; }
  // device.h:81: synchronous op execution (GPU: enqueue only).
  virtual void Compute(OpKernel* op_kernel, OpKernelContext* context) {
    op_kernel->Compute(context);
  }
  // device.h:104: blocks until all enqueued work is done.
  virtual Status Sync() = 0;
  // Makes a device-resident copy of a host tensor / host copy of a device tensor (the jobs of
  // GPUUtil::CopyCPUTensorToGPU / CopyGPUTensorToCPU, common_runtime/gpu/gpu_util.cc).
  virtual Status MakeTensorFromHost(const Tensor& host, Tensor* device_tensor) = 0;
  virtual Status CopyTensorToHost(const Tensor& device_tensor, Tensor* host) = 0;

 private:
  const std::string name_;
  const std::string device_type_;
};

}  // namespace tensorflow
#endif
