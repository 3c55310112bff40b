// Helpers shared by the B200 OpKernel wrappers: stream lookup, ABI status mapping, dtype codes.
#ifndef B200TF_CORE_KERNELS_GPU_KERNEL_UTIL_H_
#define B200TF_CORE_KERNELS_GPU_KERNEL_UTIL_H_

#include "b200_ops.h"
#include "tensorflow/core/framework/op_kernel.h"

namespace tensorflow {

// ctx->op_device_context()->stream() with the documented fallback to the device's default
// context (op_kernel.h:875-882, device_base.h:130-145).  Returns the CUstream as void*.
inline void* GetCudaStream(OpKernelContext* ctx) {
  DeviceContext* dc = ctx->op_device_context();
  if (dc == nullptr && ctx->device()->tensorflow_gpu_device_info())
    dc = ctx->device()->tensorflow_gpu_device_info()->default_context;
  if (dc == nullptr || dc->stream() == nullptr) return nullptr;
  return dc->stream()->cuda_stream();
}

// The C ABI returns tensorflow::error::Code values; b200_last_error() carries the message.
inline Status FromAbi(int rc, const char* what) {
  if (rc == 0) return Status::OK();
  return Status(static_cast<error::Code>(rc), strings::StrCat(what, ": ", b200_last_error()));
}

// Kernels are registered per dtype; this maps the template type to the ABI's DataType code.
template <typename T> struct AbiType;
template <> struct AbiType<float> { static constexpr int v = B200_DT_FLOAT; };
template <> struct AbiType<bfloat16> { static constexpr int v = B200_DT_BFLOAT16; };
template <> struct AbiType<half> { static constexpr int v = B200_DT_HALF; };
template <> struct AbiType<int32> { static constexpr int v = B200_DT_INT32; };
template <> struct AbiType<int64> { static constexpr int v = B200_DT_INT64; };

#define REGISTER_B200_FLOAT_TYPES(M) M(float) M(bfloat16)

// ---- data_format = "NCHW": the kernels of this library are NHWC-native, so an NCHW op transposes
// its activations on the way in and out (what the reference's GPU kernels do in the other
// direction around cuDNN, conv_ops.cc:558-612,712-719).
inline TensorShape NchwToNhwcShape(const TensorShape& s) {
  return TensorShape({s.dim_size(0), s.dim_size(2), s.dim_size(3), s.dim_size(1)});
}
inline TensorShape NhwcToNchwShape(const TensorShape& s) {
  return TensorShape({s.dim_size(0), s.dim_size(3), s.dim_size(1), s.dim_size(2)});
}
// nchw [N, C, H, W] -> *nhwc (allocate_temp) [N, H, W, C]
template <typename T>
Status NchwToNhwc(OpKernelContext* ctx, const Tensor& nchw, Tensor* nhwc) {
  if (nchw.dims() != 4) return errors::InvalidArgument("NCHW tensors must be 4-dimensional");
  TF_RETURN_IF_ERROR(ctx->allocate_temp(nchw.dtype(), NchwToNhwcShape(nchw.shape()), nhwc));
  return FromAbi(b200_batched_transpose(AbiType<T>::v, nchw.raw_data(), nhwc->raw_data(),
                                        nchw.dim_size(0), nchw.dim_size(1),
                                        nchw.dim_size(2) * nchw.dim_size(3), GetCudaStream(ctx)),
                 "NCHW->NHWC");
}
// nhwc [N, H, W, C] -> nchw (already allocated) [N, C, H, W]
template <typename T>
Status NhwcToNchw(OpKernelContext* ctx, const Tensor& nhwc, Tensor* nchw) {
  return FromAbi(b200_batched_transpose(AbiType<T>::v, nhwc.raw_data(), nchw->raw_data(),
                                        nhwc.dim_size(0), nhwc.dim_size(1) * nhwc.dim_size(2),
                                        nhwc.dim_size(3), GetCudaStream(ctx)),
                 "NHWC->NCHW");
}

}  // namespace tensorflow
#endif
