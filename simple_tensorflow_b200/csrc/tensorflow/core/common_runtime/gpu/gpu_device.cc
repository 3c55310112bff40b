#include "tensorflow/core/common_runtime/gpu/gpu_device.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace tensorflow {

static Status AbiStatus(int rc, const char* what) {
  if (rc == 0) return Status::OK();
  return Status(static_cast<error::Code>(rc), strings::StrCat(what, ": ", b200_last_error()));
}

void GPUDeviceContext::CopyCPUTensorToDevice(const Tensor* cpu_tensor, Device* device,
                                             Tensor* device_tensor, StatusCallback done) const {
  const size_t bytes = cpu_tensor->TotalBytes();
  if (bytes > 0) {
    gpu::DeviceMemoryBase dst(device_tensor->raw_data(), bytes);
    stream_->ThenMemcpyH2D(&dst, cpu_tensor->raw_data(), bytes);
    if (!stream_->ok()) {
      done(errors::Internal("CPU->GPU Memcpy failed: ", b200_last_error()));
      return;
    }
  }
  done(Status::OK());
}

void GPUDeviceContext::CopyDeviceTensorToCPU(const Tensor* device_tensor, const std::string&,
                                             Device* device, Tensor* cpu_tensor,
                                             StatusCallback done) {
  const size_t bytes = device_tensor->TotalBytes();
  if (bytes > 0) {
    gpu::DeviceMemoryBase src(device_tensor->raw_data(), bytes);
    stream_->ThenMemcpyD2H(cpu_tensor->raw_data(), src, bytes);
    if (!stream_->ok()) {
      done(errors::Internal("GPU->CPU Memcpy failed: ", b200_last_error()));
      return;
    }
  }
  done(Status::OK());
}

BaseGPUDevice::BaseGPUDevice(int gpu_id, const std::string& name)
    : Device(name, DEVICE_GPU), gpu_id_(gpu_id) {}

void BaseGPUDevice::set_collective_comm(void* comm, int num_replicas) {
  collective_comm_ = comm;
  num_replicas_ = num_replicas;
  const char* off = getenv("B200TF_PEER_ALLREDUCE");
  if (comm == nullptr || num_replicas < 2 || (off != nullptr && std::strcmp(off, "0") == 0)) return;
  b200_set_device(gpu_id_);
  int rank = -1;
  if (b200_nccl_comm_user_rank(comm, &rank) != 0) return;
  size_t mb = 64;
  if (const char* v = getenv("B200TF_PEER_ARENA_MB")) mb = static_cast<size_t>(std::atoll(v));
  if (b200_peer_arena_create(comm, rank, num_replicas, mb << 20, &peer_arena_) != 0)
    peer_arena_ = nullptr;  // every rank took the same decision (b200_ops.h)
}

void* BaseGPUDevice::AllocatePeerArena(size_t bytes) {
  if (peer_arena_ == nullptr) return nullptr;
  bytes = (bytes + 255) / 256 * 256;
  if (peer_arena_used_ + bytes > b200_peer_arena_bytes(peer_arena_)) return nullptr;
  char* p = static_cast<char*>(b200_peer_arena_data(peer_arena_)) + peer_arena_used_;
  peer_arena_used_ += bytes;
  return p;
}

long long BaseGPUDevice::PeerArenaOffset(const void* p) const {
  if (peer_arena_ == nullptr) return -1;
  const char* base = static_cast<const char*>(b200_peer_arena_data(peer_arena_));
  const char* q = static_cast<const char*>(p);
  if (q < base || q >= base + b200_peer_arena_bytes(peer_arena_)) return -1;
  return q - base;
}

BaseGPUDevice::~BaseGPUDevice() {
  if (peer_arena_) {
    b200_set_device(gpu_id_);
    b200_peer_arena_destroy(peer_arena_);
  }
  if (h2d_stream_) h2d_stream_->BlockHostUntilDone();
  if (collective_stream_) collective_stream_->BlockHostUntilDone();
  if (stream_) stream_->BlockHostUntilDone();
  // tensors that outlive the session keep the arena alive (gpu_bfc_allocator.h)
  if (gpu_allocator_) gpu_allocator_->Unref();
}

Status BaseGPUDevice::Create(int gpu_id, size_t memory_limit_bytes,
                             std::unique_ptr<BaseGPUDevice>* out) {
  const int n = b200_device_count();
  if (gpu_id < 0 || gpu_id >= n)
    return errors::NotFound("GPU device ", gpu_id, " not found (", n,
                            " CUDA devices visible; the B200 kernels have no CPU fallback)");
  TF_RETURN_IF_ERROR(AbiStatus(b200_set_device(gpu_id), "b200_set_device"));
  std::unique_ptr<BaseGPUDevice> d(
      new BaseGPUDevice(gpu_id, strings::StrCat("/job:localhost/replica:0/task:0/gpu:", gpu_id)));
  d->stream_.reset(new gpu::Stream());
  d->stream_->Init();
  d->h2d_stream_.reset(new gpu::Stream());
  d->h2d_stream_->Init();
  d->collective_stream_.reset(new gpu::Stream());
  {
    const char* pr = getenv("B200TF_COLLECTIVE_PRIORITY");  // default: high
    d->collective_stream_->Init(pr == nullptr || std::strcmp(pr, "0") != 0);
  }
  d->h2d_fence_.reset(new gpu::Event());
  if (!d->stream_->ok() || !d->h2d_stream_->ok() || !d->collective_stream_->ok() ||
      !d->h2d_fence_->Init())
    return errors::Internal("Failed to create the device's stream group");
  if (memory_limit_bytes == 0) {
    size_t free_b = 0, total_b = 0;
    TF_RETURN_IF_ERROR(AbiStatus(b200_mem_info(&free_b, &total_b), "b200_mem_info"));
    const size_t reserve = std::max<size_t>(300u << 20, static_cast<size_t>(free_b * 0.05));
    memory_limit_bytes = free_b > reserve ? free_b - reserve : free_b;
  }
  d->gpu_allocator_ = new GPUBFCAllocator(gpu_id, memory_limit_bytes,
                                          strings::StrCat("GPU_", gpu_id, "_bfc"));
  d->host_allocator_ = GPUHostAllocator::Process();
  d->context_.reset(new GPUDeviceContext(d->stream_.get(), d->host_allocator_));
  d->collective_context_.reset(
      new GPUDeviceContext(d->collective_stream_.get(), d->host_allocator_));
  d->gpu_device_info_.stream = d->stream_.get();
  d->gpu_device_info_.default_context = d->context_.get();
  d->gpu_device_info_.gpu_id = gpu_id;
  d->set_tensorflow_gpu_device_info(&d->gpu_device_info_);
  *out = std::move(d);
  return Status::OK();
}

void BaseGPUDevice::Compute(OpKernel* op_kernel, OpKernelContext* context) {
  // One process may drive several devices (replicas): make ours current, then enqueue.
  b200_set_device(gpu_id_);
  op_kernel->Compute(context);
  if (context->status().ok() && !stream_->ok())
    context->SetStatus(errors::Internal("GPU stream failed while running ", op_kernel->name()));
}

Status BaseGPUDevice::Sync() {
  b200_set_device(gpu_id_);
  if (!stream_->BlockHostUntilDone())
    return errors::Internal("GPU sync failed: ", b200_last_error());
  return Status::OK();
}

Status BaseGPUDevice::MakeTensorFromHost(const Tensor& host, Tensor* device_tensor) {
  b200_set_device(gpu_id_);
  Tensor t(gpu_allocator_, host.dtype(), host.shape());
  if (!t.IsInitialized())
    return errors::ResourceExhausted("OOM when allocating feed tensor ", host.shape().DebugString());
  Status s;
  context_->CopyCPUTensorToDevice(&host, this, &t, [&s](const Status& r) { s = r; });
  TF_RETURN_IF_ERROR(s);
  *device_tensor = std::move(t);
  return Status::OK();
}

Status BaseGPUDevice::StageTensorFromHost(const Tensor& host, Tensor* device_tensor,
                                          gpu::Event* ready) {
  b200_set_device(gpu_id_);
  Tensor t(gpu_allocator_, host.dtype(), host.shape());
  if (!t.IsInitialized())
    return errors::ResourceExhausted("OOM when allocating staged feed ", host.shape().DebugString());
  // The arena hands out chunks whose last use may still be queued on the compute stream.
  stream_->ThenRecordEvent(h2d_fence_.get());
  h2d_stream_->ThenWaitFor(h2d_fence_.get());
  const size_t bytes = host.TotalBytes();
  if (bytes > 0) {
    gpu::DeviceMemoryBase dst(t.raw_data(), bytes);
    h2d_stream_->ThenMemcpyH2D(&dst, host.raw_data(), bytes);
  }
  h2d_stream_->ThenRecordEvent(ready);
  if (!h2d_stream_->ok() || !stream_->ok())
    return errors::Internal("CPU->GPU staged Memcpy failed: ", b200_last_error());
  *device_tensor = std::move(t);
  return Status::OK();
}

Status BaseGPUDevice::CopyTensorToHost(const Tensor& device_tensor, Tensor* host) {
  b200_set_device(gpu_id_);
  Tensor t(host_allocator_, device_tensor.dtype(), device_tensor.shape());
  if (!t.IsInitialized())
    return errors::ResourceExhausted("OOM when allocating pinned fetch buffer");
  Status s;
  context_->CopyDeviceTensorToCPU(&device_tensor, "", this, &t, [&s](const Status& r) { s = r; });
  TF_RETURN_IF_ERROR(s);
  *host = std::move(t);
  return Status::OK();
}

}  // namespace tensorflow
