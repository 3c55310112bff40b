// Softmax / LogSoftmax / SoftmaxCrossEntropyWithLogits for DEVICE_GPU on B200.
// Checks follow SoftmaxOp::Compute (core/kernels/softmax_op.h:32-56; "Log" selected by the op
// type name prefix, :35) and SoftmaxXentWithLogitsOp::Compute (core/kernels/xent_op.cc:33-79).
#include "tensorflow/core/kernels/gpu_kernel_util.h"

namespace tensorflow {

template <typename T>
class SoftmaxOp : public OpKernel {
 public:
  explicit SoftmaxOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    log_ = def().op.compare(0, 3, "Log") == 0;
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& logits_in = ctx->input(0);
    OP_REQUIRES(ctx, TensorShapeUtils::IsMatrix(logits_in.shape()),
                errors::InvalidArgument("logits must be 2-dimensional"));
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, logits_in.shape(), &out));
    if (logits_in.NumElements() == 0) return;
    OP_REQUIRES_OK(ctx, FromAbi(b200_softmax(AbiType<T>::v, logits_in.raw_data(), out->raw_data(),
                                             logits_in.dim_size(0), logits_in.dim_size(1), log_,
                                             GetCudaStream(ctx)),
                                "Softmax"));
  }

 private:
  bool log_;
};

template <typename T, bool kScaled = false>
class SoftmaxXentWithLogitsOp : public OpKernel {
 public:
  explicit SoftmaxXentWithLogitsOp(OpKernelConstruction* ctx) : OpKernel(ctx) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& logits_in = ctx->input(0);
    const Tensor& labels_in = ctx->input(1);
    OP_REQUIRES(ctx, logits_in.IsSameSize(labels_in),
                errors::InvalidArgument("logits and labels must be same size: logits_size=",
                                        logits_in.shape().DebugString(), " labels_size=",
                                        labels_in.shape().DebugString()));
    OP_REQUIRES(ctx, TensorShapeUtils::IsMatrix(logits_in.shape()),
                errors::InvalidArgument("logits must be 2-dimensional"));
    Tensor* loss_out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({logits_in.dim_size(0)}), &loss_out));
    Tensor* back_out = nullptr;
    // (the reference forwards the logits buffer here, xent_op.cc:64-66; the fused B200 kernel
    // declares its operands __restrict__, so the output gets its own buffer)
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, logits_in.shape(), &back_out));
    if (logits_in.dim_size(0) == 0) return;
    const float* scale = nullptr;
    if (kScaled) {
      OP_REQUIRES(ctx, ctx->input(2).NumElements() == 1,
                  errors::InvalidArgument("backprop_scale must have one element"));
      scale = ctx->input(2).template data<float>();
    }
    OP_REQUIRES_OK(ctx, FromAbi(b200_softmax_xent_scaled(AbiType<T>::v, logits_in.raw_data(),
                                                         labels_in.raw_data(), loss_out->raw_data(),
                                                         back_out->raw_data(), logits_in.dim_size(0),
                                                         logits_in.dim_size(1), scale,
                                                         GetCudaStream(ctx)),
                                "SoftmaxCrossEntropyWithLogits"));
  }
};

#define REGISTER_GPU(T)                                                                       \
  REGISTER_KERNEL_BUILDER(Name("Softmax").Device(DEVICE_GPU).TypeConstraint<T>("T"),          \
                          SoftmaxOp<T>);                                                      \
  REGISTER_KERNEL_BUILDER(Name("LogSoftmax").Device(DEVICE_GPU).TypeConstraint<T>("T"),       \
                          SoftmaxOp<T>);                                                      \
  REGISTER_KERNEL_BUILDER(                                                                    \
      Name("SoftmaxCrossEntropyWithLogits").Device(DEVICE_GPU).TypeConstraint<T>("T"),        \
      SoftmaxXentWithLogitsOp<T>);
REGISTER_B200_FLOAT_TYPES(REGISTER_GPU)
#undef REGISTER_GPU
typedef SoftmaxXentWithLogitsOp<float, true> ScaledSoftmaxXentOp;
REGISTER_KERNEL_BUILDER(Name("_ScaledSoftmaxCrossEntropyWithLogits")
                            .Device(DEVICE_GPU)
                            .TypeConstraint<float>("T"),
                        ScaledSoftmaxXentOp);

}  // namespace tensorflow
