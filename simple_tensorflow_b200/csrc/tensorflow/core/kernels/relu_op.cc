// Relu / ReluGrad for DEVICE_GPU on B200.  Checks follow ReluOp / ReluGradOp
// (core/kernels/relu_op.h:34-94, numeric_op.h:52-111): same-shape inputs, in-place allowed.
#include "tensorflow/core/kernels/gpu_kernel_util.h"

namespace tensorflow {

template <typename T>
class ReluOp : public OpKernel {
 public:
  explicit ReluOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& features = ctx->input(0);
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->forward_input_or_allocate_output({0}, 0, features.shape(), &out));
    OP_REQUIRES_OK(ctx, FromAbi(b200_relu(AbiType<T>::v, features.raw_data(), out->raw_data(),
                                          features.NumElements(), GetCudaStream(ctx)),
                                "Relu"));
  }
};

template <typename T>
class ReluGradOp : public OpKernel {
 public:
  explicit ReluGradOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& g = ctx->input(0);
    const Tensor& a = ctx->input(1);
    OP_REQUIRES(ctx, a.IsSameSize(g),
                errors::InvalidArgument("Inputs must have the same size"));  // relu_op.h:48-60
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->forward_input_or_allocate_output({0, 1}, 0, g.shape(), &out));
    OP_REQUIRES_OK(ctx, FromAbi(b200_relu_grad(AbiType<T>::v, g.raw_data(), a.raw_data(),
                                               out->raw_data(), g.NumElements(),
                                               GetCudaStream(ctx)),
                                "ReluGrad"));
  }
};

// ReluGrad + BiasAddGrad(NHWC) of its result in one pass (created only by the executor's rewrite):
// outputs (backprops, bias_grad); validation = ReluGradOp's plus BiasGradOp's (bias_op.cc:185-199).
template <typename T>
class ReluGradBiasAddGradOp : public OpKernel {
 public:
  explicit ReluGradBiasAddGradOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& g = ctx->input(0);
    const Tensor& a = ctx->input(1);
    OP_REQUIRES(ctx, a.IsSameSize(g), errors::InvalidArgument("Inputs must have the same size"));
    OP_REQUIRES(ctx, TensorShapeUtils::IsMatrixOrHigher(g.shape()),
                errors::InvalidArgument("Input tensor must be at least 2D: ",
                                        g.shape().DebugString()));
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->forward_input_or_allocate_output({0}, 0, g.shape(), &out));
    const int64 channels = g.dim_size(g.dims() - 1);
    Tensor* db = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({channels}), &db));
    if (channels == 0) return;
    const int64 rows = g.NumElements() / channels;
    if (rows == 0) {  // sum over nothing (bias_op.cc:206-209 zero-fills)
      OP_REQUIRES_OK(ctx, FromAbi(b200_memset_async(db->raw_data(), 0, db->TotalBytes(),
                                                    GetCudaStream(ctx)),
                                  "BiasAddGrad"));
      return;
    }
    const size_t ws = b200_relu_grad_bias_grad_workspace_bytes(AbiType<T>::v, rows, channels);
    Tensor scratch;
    if (ws > 0)
      OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, TensorShape({static_cast<int64>(ws)}), &scratch));
    OP_REQUIRES_OK(ctx, FromAbi(b200_relu_grad_bias_grad(
                                    AbiType<T>::v, g.raw_data(), a.raw_data(), out->raw_data(),
                                    db->raw_data(), rows, channels,
                                    ws ? scratch.raw_data() : nullptr, ws, GetCudaStream(ctx)),
                                "_ReluGradBiasAddGrad"));
  }
};

#define REGISTER_GPU(T)                                                                      \
  REGISTER_KERNEL_BUILDER(Name("_ReluGradBiasAddGrad").Device(DEVICE_GPU).TypeConstraint<T>("T"), \
                          ReluGradBiasAddGradOp<T>);                                         \
  REGISTER_KERNEL_BUILDER(Name("Relu").Device(DEVICE_GPU).TypeConstraint<T>("T"), ReluOp<T>); \
  REGISTER_KERNEL_BUILDER(Name("ReluGrad").Device(DEVICE_GPU).TypeConstraint<T>("T"),        \
                          ReluGradOp<T>);
REGISTER_B200_FLOAT_TYPES(REGISTER_GPU)
#undef REGISTER_GPU

}  // namespace tensorflow
