// BaseGPUDevice for B200 -- the "new tensorflow/core/common_runtime/gpu device" of the
// north star.  Same roles as the reference's core/common_runtime/gpu/gpu_device.{h,cc}:
//   * a stream group per device (BaseGPUDevice::Init, gpu_device.cc:194-264): ONE compute stream
//     that every kernel runs on, a host_to_device stream for staged feeds, and (additive) a
//     collective stream so gradient all-reduces overlap the rest of the backward pass
//   * BFC arena for device memory, pinned-host allocator for feeds/fetches
//   * GPUDeviceContext carrying the stream to kernels (gpu_device_context.h) and doing the
//     H2D / D2H tensor copies of GPUUtil (gpu_util.cc)
//   * Compute() only enqueues (gpu_device.cc:337-399); Sync() is the one blocking call per step
//     (gpu_device.cc:413)
// All CUDA access goes through the C ABI shim (tensorflow/stream_executor/stream.h).
#ifndef B200TF_CORE_COMMON_RUNTIME_GPU_GPU_DEVICE_H_
#define B200TF_CORE_COMMON_RUNTIME_GPU_GPU_DEVICE_H_

#include <memory>
#include <string>

#include "tensorflow/core/common_runtime/device.h"
#include "tensorflow/core/common_runtime/gpu/gpu_bfc_allocator.h"

namespace tensorflow {

class GPUDeviceContext : public DeviceContext {
 public:
  GPUDeviceContext(gpu::Stream* stream, Allocator* host_allocator)
      : stream_(stream), host_allocator_(host_allocator) {}
  gpu::Stream* stream() const override { return stream_; }
  void CopyCPUTensorToDevice(const Tensor* cpu_tensor, Device* device, Tensor* device_tensor,
                             StatusCallback done) const override;
  void CopyDeviceTensorToCPU(const Tensor* device_tensor, const std::string& tensor_name,
                             Device* device, Tensor* cpu_tensor, StatusCallback done) override;

 private:
  gpu::Stream* stream_;
  Allocator* host_allocator_;
};

class BaseGPUDevice : public Device {
 public:
  // memory_limit_bytes == 0: free memory minus max(300 MiB, 5 %) like gpu_device.cc:546-561.
  static Status Create(int gpu_id, size_t memory_limit_bytes, std::unique_ptr<BaseGPUDevice>* out);
  ~BaseGPUDevice() override;

  Allocator* GetAllocator(AllocatorAttributes attr) override {
    return attr.on_host() ? static_cast<Allocator*>(host_allocator_)
                          : static_cast<Allocator*>(gpu_allocator_);
  }
  void Compute(OpKernel* op_kernel, OpKernelContext* context) override;
  Status Sync() override;
  Status MakeTensorFromHost(const Tensor& host, Tensor* device_tensor) override;
  Status CopyTensorToHost(const Tensor& device_tensor, Tensor* host) override;

  // Replica data-parallel: the NCCL communicator this device's B200AllReduce kernels use
  // (created by the host with b200_nccl_comm_init_rank; not owned).
  // Also maps an NVLink peer arena across the replicas (collective over `comm`; skipped with
  // B200TF_PEER_ALLREDUCE=0 or when any rank cannot map its peers -> NCCL carries the exchange).
  void set_collective_comm(void* comm, int num_replicas);
  // The peer arena: this step's gradient arenas are carved from it front to back.
  void* peer_arena() const { return peer_arena_; }
  void ResetPeerArena() { peer_arena_used_ = 0; }
  // Reserves `bytes` (256-byte aligned) of the peer arena; nullptr when it does not fit.
  void* AllocatePeerArena(size_t bytes);
  // Byte offset of `p` inside the peer arena's data region, or -1 when it is not part of it.
  long long PeerArenaOffset(const void* p) const;
  void* collective_comm() const { return collective_comm_; }
  int num_replicas() const { return num_replicas_; }

  int gpu_id() const { return gpu_id_; }
  gpu::Stream* compute_stream() const { return stream_.get(); }
  gpu::Stream* host_to_device_stream() const { return h2d_stream_.get(); }
  gpu::Stream* collective_stream() const { return collective_stream_.get(); }
  GPUDeviceContext* device_context() const { return context_.get(); }
  // The DeviceContext handed to kernels the executor places on the collective stream.
  GPUDeviceContext* collective_context() const { return collective_context_.get(); }

  // Starts the host->device copy of `host` on the host_to_device stream (after everything the
  // compute stream has been given so far, like GPUUtil::CopyCPUTensorToGPU's
  // ThenWaitFor(compute), gpu_util.cc:300-306) and returns at once; `ready` is recorded behind
  // the copy.  The caller makes the consuming stream wait for `ready`.
  Status StageTensorFromHost(const Tensor& host, Tensor* device_tensor, gpu::Event* ready);
  Allocator* host_allocator() const { return host_allocator_; }

 private:
  BaseGPUDevice(int gpu_id, const std::string& name);
  const int gpu_id_;
  std::unique_ptr<gpu::Stream> stream_, h2d_stream_, collective_stream_;
  std::unique_ptr<gpu::Event> h2d_fence_;
  GPUBFCAllocator* gpu_allocator_ = nullptr;   // ref-counted: Unref() in the destructor
  GPUHostAllocator* host_allocator_ = nullptr;  // GPUHostAllocator::Process(), never destroyed
  std::unique_ptr<GPUDeviceContext> context_, collective_context_;
  GpuDeviceInfo gpu_device_info_;
  void* collective_comm_ = nullptr;
  int num_replicas_ = 1;
  void* peer_arena_ = nullptr;
  size_t peer_arena_used_ = 0;
};

}  // namespace tensorflow
#endif
