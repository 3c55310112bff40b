// BiasAdd / BiasAddGrad for DEVICE_GPU on B200.  Both layouts are native: NHWC streams [rows, C];
// data_format NCHW (GPU-only in the reference too, bias_op.cc:242-299) streams [batch, C, H * W]
// planes through b200_bias_add_nchw / b200_bias_add_grad_nchw, with the reference's dimension
// rule (GetBiasValueDims, bias_op.cc:127-151: channel = dims - 3, batch = everything before).
// Validation follows BiasOp::Compute (core/kernels/bias_op.cc:62-117) and BiasGradOp::Compute
// (:185-227); launches replace BiasGPU / BiasGradGPU (bias_op_gpu.cu.cc:69-88,189-242).
#include "tensorflow/core/kernels/gpu_kernel_util.h"
#include "tensorflow/core/util/padding.h"

namespace tensorflow {

static Status ReadFormat(OpKernelConstruction* ctx, bool* nchw) {
  std::string data_format;
  *nchw = false;
  if (ctx->GetAttr("data_format", &data_format).ok()) {
    TensorFormat fmt;
    if (!FormatFromString(data_format, &fmt)) return errors::InvalidArgument("Invalid data format");
    *nchw = fmt == FORMAT_NCHW;
  }
  return Status::OK();
}

// GetBiasValueDims for FORMAT_NCHW (bias_op.cc:140-150): [batch dims..., C, H, W]
static void NchwDims(const Tensor& t, int64* batch, int64* channels, int64* image) {
  const int cd = t.dims() - 3;
  *batch = 1;
  for (int i = 0; i < cd; ++i) *batch *= t.dim_size(i);
  *channels = t.dim_size(cd);
  *image = t.dim_size(cd + 1) * t.dim_size(cd + 2);
}

template <typename T>
class BiasOp : public OpKernel {
 public:
  explicit BiasOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ReadFormat(ctx, &nchw_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& input = ctx->input(0);
    const Tensor& bias = ctx->input(1);
    OP_REQUIRES(ctx, TensorShapeUtils::IsMatrixOrHigher(input.shape()),
                errors::InvalidArgument("Input tensor must be at least 2D: ",
                                        input.shape().DebugString()));
    OP_REQUIRES(ctx, TensorShapeUtils::IsVector(bias.shape()),
                errors::InvalidArgument("Biases must be 1D: ", bias.shape().DebugString()));
    if (nchw_ && input.dims() > 2) {
      int64 batch, channels, image;
      NchwDims(input, &batch, &channels, &image);
      OP_REQUIRES(ctx, bias.dim_size(0) == channels,
                  errors::InvalidArgument("Must provide as many biases as the channel dimension "
                                          "of the input tensor: ", bias.shape().DebugString(),
                                          " vs. ", channels, " in ", input.shape().DebugString()));
      Tensor* output = nullptr;
      OP_REQUIRES_OK(ctx, ctx->forward_input_or_allocate_output({0}, 0, input.shape(), &output));
      if (input.NumElements() == 0) return;
      OP_REQUIRES_OK(ctx, FromAbi(b200_bias_add_nchw(AbiType<T>::v, input.raw_data(),
                                                     bias.raw_data(), output->raw_data(), batch,
                                                     channels, image, GetCudaStream(ctx)),
                                  "BiasAdd"));
      return;
    }
    const int64 channels = input.dim_size(input.dims() - 1);
    OP_REQUIRES(ctx, bias.dim_size(0) == channels,
                errors::InvalidArgument("Must provide as many biases as the last dimension of the "
                                        "input tensor: ", bias.shape().DebugString(), " vs. ",
                                        input.shape().DebugString()));
    Tensor* output = nullptr;
    OP_REQUIRES_OK(ctx, ctx->forward_input_or_allocate_output({0}, 0, input.shape(), &output));
    if (input.NumElements() == 0) return;
    OP_REQUIRES_OK(ctx, FromAbi(b200_bias_add(AbiType<T>::v, input.raw_data(), bias.raw_data(),
                                              output->raw_data(), input.NumElements() / channels,
                                              channels, GetCudaStream(ctx)),
                                "BiasAdd"));
  }

 private:
  bool nchw_;
};

template <typename T>
class BiasGradOp : public OpKernel {
 public:
  explicit BiasGradOp(OpKernelConstruction* ctx) : OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ReadFormat(ctx, &nchw_));
  }
  void Compute(OpKernelContext* ctx) override {
    Tensor g = ctx->input(0);
    OP_REQUIRES(ctx, TensorShapeUtils::IsMatrixOrHigher(g.shape()),
                errors::InvalidArgument("Input tensor must be at least 2D: ",
                                        g.shape().DebugString()));
    if (nchw_ && g.dims() > 2) {
      int64 batch, channels, image;
      NchwDims(g, &batch, &channels, &image);
      Tensor* output = nullptr;
      OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({channels}), &output));
      if (channels == 0) return;
      const size_t ws = b200_bias_add_grad_nchw_workspace_bytes(AbiType<T>::v, batch, channels, image);
      Tensor scratch;
      if (ws > 0)
        OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, TensorShape({(int64)ws}), &scratch));
      OP_REQUIRES_OK(ctx, FromAbi(b200_bias_add_grad_nchw(AbiType<T>::v, g.raw_data(),
                                                          output->raw_data(), batch, channels,
                                                          image, ws ? scratch.raw_data() : nullptr,
                                                          ws, GetCudaStream(ctx)),
                                  "BiasAddGrad"));
      return;
    }
    const int64 channels = g.dim_size(g.dims() - 1);
    Tensor* output = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({channels}), &output));
    if (channels == 0) return;
    const int64 rows = g.NumElements() / channels;
    const size_t ws = b200_bias_add_grad_workspace_bytes(AbiType<T>::v, rows, channels);
    Tensor scratch;
    if (ws > 0) OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, TensorShape({(int64)ws}), &scratch));
    OP_REQUIRES_OK(ctx, FromAbi(b200_bias_add_grad(AbiType<T>::v, g.raw_data(), output->raw_data(),
                                                   rows, channels, ws ? scratch.raw_data() : nullptr,
                                                   ws, GetCudaStream(ctx)),
                                "BiasAddGrad"));
  }

 private:
  bool nchw_;
};

#define REGISTER_GPU(T)                                                                     \
  REGISTER_KERNEL_BUILDER(Name("BiasAdd").Device(DEVICE_GPU).TypeConstraint<T>("T"),        \
                          BiasOp<T>);                                                       \
  REGISTER_KERNEL_BUILDER(Name("BiasAddGrad").Device(DEVICE_GPU).TypeConstraint<T>("T"),    \
                          BiasGradOp<T>);
REGISTER_B200_FLOAT_TYPES(REGISTER_GPU)
#undef REGISTER_GPU

}  // namespace tensorflow
